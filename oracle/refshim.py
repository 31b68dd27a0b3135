"""ORACLE (test infrastructure, not product code).

Import the UNMODIFIED reference modules from /root/reference (present only in
the build container, never on the GPU box) so the oracle restatements can be
pinned against them and golden fixtures generated (oracle/make_golden.py).

Shims needed because of missing third-party packages:
  * ``munkres``            -> oracle/munkres_ref.py (faithful restatement; NOT scipy)
  * ``dataset.transforms`` -> stub module exposing FLIP_CONFIG
    (literal of reference lib/dataset/transforms/build.py:15-28) so that
    lib/core/inference.py imports without pulling pycocotools.
Nothing from the reference is copied; modules are imported in place.
"""
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("LITEPOSE_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "models"))


_loaded = {}


def load():
    """Returns a namespace with .pose_mobilenet, .group, .inference, .fp16util."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    from . import munkres_ref
    from litepose_b200.config import FLIP_CONFIG

    # the reference's own import roots (cf. reference _init_paths.py:20-23)
    for p in (REF_ROOT, os.path.join(REF_ROOT, "lib")):
        if p not in sys.path:
            sys.path.append(p)
    saved = {k: sys.modules.get(k) for k in ("munkres", "dataset", "dataset.transforms")}
    mk = types.ModuleType("munkres")
    mk.Munkres = munkres_ref.Munkres
    sys.modules["munkres"] = mk
    ds = types.ModuleType("dataset")
    dst = types.ModuleType("dataset.transforms")
    dst.FLIP_CONFIG = FLIP_CONFIG
    ds.transforms = dst
    sys.modules["dataset"] = ds
    sys.modules["dataset.transforms"] = dst
    import importlib
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pm = importlib.import_module("lib.models.pose_mobilenet")
        grp = importlib.import_module("lib.core.group")
        inf = importlib.import_module("lib.core.inference")
        f16 = importlib.import_module("lib.fp16_utils.fp16util")
    ns = types.SimpleNamespace(pose_mobilenet=pm, group=grp, inference=inf, fp16util=f16)
    _loaded["ns"] = ns
    return ns


def build_reference_model(cfg, arch, seed=0, bn_seed=1):
    """Reference LitePose with the synthetic weights of SURVEY §8d."""
    import torch
    from litepose_b200.synth import randomize_bn_
    ns = load()
    torch.manual_seed(seed)
    model = ns.pose_mobilenet.get_pose_net(cfg, False, arch)
    randomize_bn_(model, bn_seed)
    return model.eval()
