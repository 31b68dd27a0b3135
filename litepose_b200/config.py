"""Config objects for the LitePose inference path.

The reference drives everything from a yacs ``CfgNode`` (reference
lib/config/default.py:20-153 defaults, overridden by
experiments/crowd_pose/mobilenet/mobile.yaml and ``--superconfig`` JSON,
valid.py:103-111).  yacs is not part of this image, and the hot path only reads
attributes, so this module provides an attribute-dict with the same field names
and the values of the evaluation config of record (mobile.yaml).  A real yacs
``cfg`` works unchanged everywhere a ``cfg`` is accepted.
"""
import copy
import json
import os


class CfgNode(dict):
    """Minimal attribute-access dict (read/write), nestable, deep-copyable."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)


def _cn(d):
    out = CfgNode()
    for k, v in d.items():
        out[k] = _cn(v) if isinstance(v, dict) else v
    return out


# The four searched architectures shipped by the reference
# (mobile_configs/search-{XS,S,M,L}.json).  They are data, reproduced here as the
# generating rule so the package does not depend on /root/reference at run time:
# every stage is ``num_blocks`` InvBottlenecks with expansion 6 and 7x7 depthwise.
_ARCH_TABLE = {
    #        img  c0  deconv           stage channels     blocks        strides
    "XS": (256, 16, [16, 24, 24], [16, 32, 48, 80], [6, 8, 10, 10], [2, 2, 2, 1]),
    "S": (448, 16, [32, 24, 32], [16, 32, 48, 120], [6, 8, 10, 10], [2, 2, 2, 1]),
    "M": (448, 16, [64, 40, 32], [24, 48, 72, 120], [6, 8, 10, 10], [2, 2, 2, 1]),
    "L": (512, 24, [64, 40, 32], [24, 64, 96, 160], [6, 8, 10, 10], [2, 2, 2, 1]),
}


def get_arch(name_or_path):
    """Return a ``cfg_arch`` dict: either one of 'XS','S','M','L' or a JSON path
    in the reference's mobile_configs format."""
    if isinstance(name_or_path, dict):
        return name_or_path
    if name_or_path in _ARCH_TABLE:
        img, c0, dec, chans, blocks, strides = _ARCH_TABLE[name_or_path]
        return {
            "img_size": img,
            "input_channel": c0,
            "deconv_setting": list(dec),
            "backbone_setting": [
                {"num_blocks": n, "stride": s, "channel": c,
                 "block_setting": [[6, 7] for _ in range(n)]}
                for c, n, s in zip(chans, blocks, strides)
            ],
        }
    if os.path.isfile(name_or_path):
        with open(name_or_path, "r") as f:
            return json.load(f)
    raise ValueError("unknown architecture: %r" % (name_or_path,))


def get_cfg(dataset="crowd_pose", input_size=512, flip_test=True, project2image=True,
            adjust=True, refine=True):
    """cfg with the values of experiments/crowd_pose/mobilenet/mobile.yaml on top of
    lib/config/default.py defaults (only the keys the inference path reads)."""
    nj = 14 if dataset == "crowd_pose" else 17
    cfg = _cn({
        "GPUS": (0,),
        "FP16": {"ENABLED": True},
        "CUDNN": {"BENCHMARK": True, "DETERMINISTIC": False, "ENABLED": True},
        "MODEL": {
            "NAME": "pose_mobilenet",
            "INIT_WEIGHTS": False,
            "PRETRAINED": "",
            "NUM_JOINTS": nj,
            "TAG_PER_JOINT": True,
            "EXTRA": {
                "FINAL_CONV_KERNEL": 1,
                "NUM_DECONV_LAYERS": 3,
                "NUM_DECONV_FILTERS": [64, 48, 32],
                "NUM_DECONV_KERNELS": [4, 4, 4],
            },
        },
        "LOSS": {
            "NUM_STAGES": 2,
            "WITH_HEATMAPS_LOSS": (True, True),
            "WITH_AE_LOSS": (True, False),
        },
        "DATASET": {
            "DATASET": "crowd_pose_kpt" if dataset == "crowd_pose" else "coco_kpt",
            "DATASET_TEST": dataset,
            "NUM_JOINTS": nj,
            "MAX_NUM_PEOPLE": 30,
            "INPUT_SIZE": input_size,
            "OUTPUT_SIZE": [input_size // 4, input_size // 2],
            "WITH_CENTER": False,
        },
        "TEST": {
            "FLIP_TEST": flip_test,
            "ADJUST": adjust,
            "REFINE": refine,
            "SCALE_FACTOR": [1],
            "DETECTION_THRESHOLD": 0.1,
            "TAG_THRESHOLD": 1.0,
            "USE_DETECTION_VAL": True,
            "IGNORE_TOO_MUCH": False,
            "IGNORE_CENTER": True,
            "NMS_KERNEL": 5,
            "NMS_PADDING": 2,
            "PROJECT2IMAGE": project2image,
            "WITH_HEATMAPS": (True, True),
            "WITH_AE": (True, False),
            "IMAGES_PER_GPU": 1,
        },
    })
    return cfg


# lib/dataset/transforms/build.py:15-28 -- channel permutation applied to the
# outputs of the horizontally flipped pass.
FLIP_CONFIG = {
    "COCO": [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15],
    "COCO_WITH_CENTER": [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15, 17],
    "CROWDPOSE": [1, 0, 3, 2, 5, 4, 7, 6, 9, 8, 11, 10, 12, 13],
    "CROWDPOSE_WITH_CENTER": [1, 0, 3, 2, 5, 4, 7, 6, 9, 8, 11, 10, 12, 13, 14],
}


def flip_index_for(cfg):
    """Same selection rule as lib/core/inference.py:108-117."""
    if "coco" in cfg.DATASET.DATASET:
        name = "COCO"
    elif "crowd_pose" in cfg.DATASET.DATASET:
        name = "CROWDPOSE"
    else:
        raise ValueError("Please implement flip_index for new dataset: %s." % cfg.DATASET.DATASET)
    if cfg.DATASET.WITH_CENTER:
        name += "_WITH_CENTER"
    return FLIP_CONFIG[name]


# ---- yacs-free loader for the reference's experiment files ------------------------------------------------------------
# Defaults of the keys the inference path reads, as lib/config/default.py:20-153 sets them (the training / debug
# sections of an experiment file are merged as they come: the reference's root node allows new keys).
_DEFAULTS = {
    "GPUS": (0,), "WORKERS": 4, "PRINT_FREQ": 20, "DATA_DIR": "", "OUTPUT_DIR": "", "LOG_DIR": "",
    "FP16": {"ENABLED": False, "STATIC_LOSS_SCALE": 1.0, "DYNAMIC_LOSS_SCALE": False},
    "CUDNN": {"BENCHMARK": True, "DETERMINISTIC": False, "ENABLED": True},
    "MODEL": {"NAME": "pose_multi_resolution_net_v16", "INIT_WEIGHTS": True, "PRETRAINED": "", "NUM_JOINTS": 17,
              "TAG_PER_JOINT": True, "EXTRA": {}, "SYNC_BN": False},
    "LOSS": {"NUM_STAGES": 1, "WITH_HEATMAPS_LOSS": (True,), "HEATMAPS_LOSS_FACTOR": (1.0,), "WITH_AE_LOSS": (True,),
             "AE_LOSS_TYPE": "max", "PUSH_LOSS_FACTOR": (0.001,), "PULL_LOSS_FACTOR": (0.001,)},
    "DATASET": {"ROOT": "", "DATASET": "coco_kpt", "DATASET_TEST": "coco", "NUM_JOINTS": 17, "MAX_NUM_PEOPLE": 30,
                "INPUT_SIZE": 512, "OUTPUT_SIZE": [128, 256, 512], "WITH_CENTER": False},
    "TEST": {"IMAGES_PER_GPU": 32, "FLIP_TEST": False, "ADJUST": True, "REFINE": True, "SCALE_FACTOR": [1],
             "DETECTION_THRESHOLD": 0.2, "TAG_THRESHOLD": 1.0, "USE_DETECTION_VAL": True, "IGNORE_TOO_MUCH": False,
             "MODEL_FILE": "", "IGNORE_CENTER": True, "NMS_KERNEL": 3, "NMS_PADDING": 1, "PROJECT2IMAGE": False,
             "WITH_HEATMAPS": (True,), "WITH_AE": (True,), "LOG_PROGRESS": False},
}


def _decode(v):
    """yacs decodes string leaves with ast.literal_eval where that succeeds ("(True, False)" -> tuple, "1e-4" -> float)."""
    import ast
    if isinstance(v, dict):
        return {k: _decode(x) for k, x in v.items()}
    if isinstance(v, str):
        try:
            return ast.literal_eval(v)
        except (ValueError, SyntaxError):
            return v
    return v


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)


def load_experiment(yaml_path, superconfig=None, opts=()):
    """What the reference's scripts do with ``--cfg FILE [--superconfig ARCH.json] [KEY VALUE ...]`` (valid.py:95-111,
    lib/config/default.py:156-190), without yacs: defaults <- experiment file <- ``opts`` pairs, the WITH_CENTER joint
    count adjustment, then the architecture's ``img_size`` as INPUT_SIZE / OUTPUT_SIZE.  Returns (cfg, cfg_arch);
    cfg_arch is None without ``superconfig`` (a path, an arch name of get_arch, or a dict)."""
    import yaml
    tree = copy.deepcopy(_DEFAULTS)
    with open(yaml_path, "r") as f:
        _merge(tree, _decode(yaml.safe_load(f) or {}))
    opts = list(opts)
    if len(opts) % 2:
        raise ValueError("opts must be KEY VALUE pairs")
    for key, val in zip(opts[0::2], opts[1::2]):
        node = tree
        parts = key.split(".")
        for p in parts[:-1]:
            node = node[p]
        if parts[-1] not in node:
            raise KeyError("unknown config key %s" % key)
        node[parts[-1]] = _decode(val)
    ds, mdl, loss = tree["DATASET"], tree["MODEL"], tree["LOSS"]
    if ds["WITH_CENTER"]:
        ds["NUM_JOINTS"] += 1
        mdl["NUM_JOINTS"] = ds["NUM_JOINTS"]
    if not isinstance(ds["OUTPUT_SIZE"], (list, tuple)):
        ds["OUTPUT_SIZE"] = [ds["OUTPUT_SIZE"]]
    for k in ("WITH_HEATMAPS_LOSS", "HEATMAPS_LOSS_FACTOR"):
        if not isinstance(loss[k], (list, tuple)):
            loss[k] = (loss[k],)
    arch = None
    if superconfig is not None:
        arch = get_arch(superconfig)
        reso = int(arch["img_size"])
        ds["INPUT_SIZE"] = reso
        ds["OUTPUT_SIZE"] = [reso // 4, reso // 2]
    return _cn(tree), arch
