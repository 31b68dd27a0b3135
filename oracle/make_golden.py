"""ORACLE (test infrastructure, not product code).

Generate the committed known-answer fixtures under tests/golden/ by running the
UNMODIFIED reference (imported in place from /root/reference through
oracle/refshim.py) on seeded synthetic inputs.  The reference ships no golden
vectors or tests of its own, so these are the pins for both the
oracle restatements and the CUDA path.  Run in the build container only:

    python -m oracle.make_golden
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from litepose_b200 import synth  # noqa: E402
from litepose_b200.config import get_arch, get_cfg  # noqa: E402
from oracle import refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

TINY_ARCH = {
    "img_size": 64, "input_channel": 16, "deconv_setting": [16, 24, 24],
    "backbone_setting": [
        {"num_blocks": 2, "stride": 2, "channel": 16, "block_setting": [[6, 7], [6, 7]]},
        {"num_blocks": 2, "stride": 2, "channel": 24, "block_setting": [[6, 7], [6, 7]]},
        {"num_blocks": 2, "stride": 2, "channel": 40, "block_setting": [[6, 7], [6, 7]]},
        {"num_blocks": 2, "stride": 1, "channel": 48, "block_setting": [[6, 7], [6, 7]]},
    ],
}


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def golden_model():
    ns = refshim.load()
    # (1) tiny arch: weights + per-layer activations committed
    cfg = get_cfg(input_size=64)
    model = refshim.build_reference_model(cfg, TINY_ARCH)
    sd = model.state_dict()
    x = synth.make_frames(2, 64, seed=7)
    acts = {}
    hooks = []
    names = ["first.0", "first.1", "first", "stage.0.0.inv", "stage.0.0.depth_conv", "stage.0.0",
             "stage.0.1", "stage.1.0", "stage.2.1", "stage.3.1", "deconv_bnrelu.0", "deconv_bnrelu.1",
             "deconv_bnrelu.2", "final_refined.0.conv.2", "final_raw.1.conv.2"]
    mods = dict(model.named_modules())
    for nme in names:
        hooks.append(mods[nme].register_forward_hook(
            lambda m, i, o, nme=nme: acts.__setitem__(nme, o.detach().clone())))
    with torch.no_grad():
        outs = model(x)
    for h in hooks:
        h.remove()
    blob = {"x": x.numpy(), "out0": outs[0].numpy(), "out1": outs[1].numpy()}
    for k, v in sd.items():
        blob["sd/" + k] = v.numpy()
    for k, v in acts.items():
        blob["act/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "model_tiny.npz"), **blob)

    # (2) shipped archs with seeded weights: digest + outputs only
    for name, size in (("XS", 128), ("S", 128)):
        cfg = get_cfg(input_size=size)
        arch = get_arch(name)
        model = refshim.build_reference_model(cfg, arch)
        x = synth.make_frames(1, size, seed=11)
        with torch.no_grad():
            outs = model(x)
            half = ns.fp16util.network_to_half(refshim.build_reference_model(cfg, arch))
            outs_h = half(x)
        np.savez_compressed(
            os.path.join(OUT, "model_%s_%d.npz" % (name, size)),
            digest=np.array(sd_digest(model.state_dict())),
            out0=outs[0].numpy(), out1=outs[1].numpy(),
            out0_fp16eager=outs_h[0].numpy(), out1_fp16eager=outs_h[1].numpy())


def golden_glue():
    ns = refshim.load()
    g = torch.Generator().manual_seed(21)
    for flip, proj in ((True, True), (False, True), (True, False)):
        cfg = get_cfg(input_size=64, flip_test=flip, project2image=proj)
        a0 = torch.randn(2, 28, 16, 16, generator=g)
        a1 = torch.randn(2, 14, 32, 32, generator=g)
        b0 = torch.randn(2, 28, 16, 16, generator=g)
        b1 = torch.randn(2, 14, 32, 32, generator=g)
        calls = []

        def fake_model(img):
            calls.append(1)
            return [a0.clone(), a1.clone()] if len(calls) == 1 else [b0.clone(), b1.clone()]

        img = torch.zeros(2, 3, 64, 64)
        _, h, t = ns.inference.get_multi_stage_outputs(cfg, fake_model, img, flip, proj, (64, 64))
        fh, tl = ns.inference.aggregate_results(cfg, 1, None, [], h, t)
        tags = torch.cat(tl, dim=4)
        np.savez_compressed(os.path.join(OUT, "glue_flip%d_proj%d.npz" % (flip, proj)),
                            a0=a0.numpy(), a1=a1.numpy(), b0=b0.numpy(), b1=b1.numpy(),
                            final_heatmaps=fh.numpy(), tags=tags.numpy())


PARSER_CASES = [
    # name, J-dataset, h, w, T, people, seed
    ("p5_128_t2", "crowd_pose", 128, 128, 2, 5, 0),
    ("p5_128_t1", "crowd_pose", 128, 128, 1, 5, 1),
    ("p30_256_t2", "crowd_pose", 256, 256, 2, 30, 2),
    ("p3_64x96_t2", "crowd_pose", 64, 96, 2, 3, 3),
    ("p0_128_t2", "crowd_pose", 128, 128, 2, 0, 4),
    ("p12_256_t2", "crowd_pose", 256, 256, 2, 12, 5),
    ("p30_256_t1", "crowd_pose", 256, 256, 1, 30, 6),
    ("coco_p8_128_t2", "coco", 128, 160, 2, 8, 7),
    # the bench geometries (BASELINE configs 3 and 5): 512^2 with 5 persons, 640^2 with a 30-person crowd
    ("p5_512_t2", "crowd_pose", 512, 512, 2, 5, 8),
    ("p30_640_t2", "crowd_pose", 640, 640, 2, 30, 9),
]


def golden_parser(only=None):
    ns = refshim.load()
    for name, ds, h, w, t, people, seed in PARSER_CASES:
        if only and name not in only:
            continue
        cfg = get_cfg(dataset=ds, input_size=256)
        nj = cfg.DATASET.NUM_JOINTS
        det, tag = synth.plant_crowd(nj, h, w, t, num_people=people, seed=seed)
        dt = torch.from_numpy(det)[None]
        tt = torch.from_numpy(tag)[None]
        rp = ns.group.HeatmapParser(cfg)
        top = rp.top_k(dt, tt)
        blob = {"in_digest": np.array(hashlib.sha256(det.tobytes() + tag.tobytes()).hexdigest()),
                "val_k": top["val_k"], "loc_k": top["loc_k"], "tag_k": top["tag_k"]}
        for adj, ref in ((True, True), (True, False), (False, False)):
            ans, scores = rp.parse(dt.clone(), tt.clone(), adj, ref)
            a = np.array(ans[0], dtype=np.float32).reshape(-1, nj, 3 + t)
            blob["ans_a%d_r%d" % (adj, ref)] = a
            blob["scores_a%d_r%d" % (adj, ref)] = np.array(scores, dtype=np.float32)
        np.savez_compressed(os.path.join(OUT, "parser_%s.npz" % name), **blob)


def golden_munkres():
    """Self-pinned restatement outputs on degenerate matrices (parity unpinned vs
    PyPI munkres, see oracle/munkres_ref.py) + the reference's own cost recipe."""
    from oracle.munkres_ref import Munkres
    rng = np.random.RandomState(5)
    mats, sols = [], []
    for trial in range(40):
        r = rng.randint(1, 13)
        c = rng.randint(1, 13)
        dist = rng.choice([0.0, 0.0, 0.0, 1.0, 2.0, 4.0], size=(r, c))
        val = rng.uniform(0.1, 1.0, size=(r, 1)).astype(np.float32).astype(np.float64)
        m = np.round(dist) * 100 - val
        if r > c:
            m = np.concatenate((m, np.zeros((r, r - c)) + 1e10), axis=1)
        pairs = Munkres().compute(m.copy())
        full = np.full((12, 12), np.nan)
        full[:m.shape[0], :m.shape[1]] = m
        sol = np.full((12, 2), -1, np.int32)
        sol[:len(pairs)] = np.array(pairs, np.int32).reshape(-1, 2)
        mats.append(full)
        sols.append(sol)
    np.savez_compressed(os.path.join(OUT, "munkres_cases.npz"), mats=np.stack(mats), sols=np.stack(sols))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 2 and sys.argv[1] == "--parser-only":       # regenerate selected parser cases only
        golden_parser(only=set(sys.argv[2:]))
    else:
        golden_model()
        golden_glue()
        golden_parser()
        golden_munkres()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
