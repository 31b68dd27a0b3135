"""CPU: the oracle restatements against the committed golden vectors that were
produced by the unmodified reference (oracle/make_golden.py)."""
import glob
import hashlib
import os

import numpy as np
import pytest

from parity_util import assert_topk_equal
import torch

from litepose_b200 import synth
from litepose_b200.config import get_arch, get_cfg
from oracle import glue_ref, group_ref, model_ref
from oracle.make_golden import PARSER_CASES, TINY_ARCH


def test_model_tiny_layers_and_outputs(golden_dir):
    z = np.load(os.path.join(golden_dir, "model_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    x = torch.from_numpy(z["x"])
    cap = {}
    with torch.no_grad():
        outs = model_ref.forward(sd, TINY_ARCH, x, cap)
    assert np.array_equal(outs[0].numpy(), z["out0"]) or np.allclose(outs[0].numpy(), z["out0"], atol=1e-6)
    assert np.allclose(outs[1].numpy(), z["out1"], atol=1e-6)
    pairs = {"first.0": "first.0", "first.1": "first.1", "first": "first",
             "stage.0.0.inv": "stage.0.0.inv", "stage.0.0.depth_conv": "stage.0.0.depth_conv",
             "stage.0.0": "stage.0.0.out", "stage.0.1": "stage.0.1.out", "stage.1.0": "stage.1.0.out",
             "stage.2.1": "stage.2.1.out", "stage.3.1": "stage.3.1.out",
             "deconv_bnrelu.0": "deconv.0", "deconv_bnrelu.1": "deconv.1", "deconv_bnrelu.2": "deconv.2",
             "final_refined.0.conv.2": "final_refined.0.conv.dw", "final_raw.1.conv.2": "final_raw.1.conv.dw"}
    for ref_name, my_name in pairs.items():
        assert np.allclose(cap[my_name].numpy(), z["act/" + ref_name], atol=1e-5), ref_name
    # BN fold parity (pattern of reference fuse_bn.py:205-216)
    with torch.no_grad():
        fold = model_ref.forward_folded(model_ref.fold_bn(sd, TINY_ARCH), TINY_ARCH, x)
    for a, k in zip(fold, ("out0", "out1")):
        assert np.abs(a.numpy() - z[k]).max() <= 1e-5 * max(1.0, np.abs(z[k]).max())


@pytest.mark.parametrize("name,size", [("XS", 128), ("S", 128)])
def test_model_shipped_arch_outputs(golden_dir, name, size):
    from litepose_b200.lib.models.pose_mobilenet import get_pose_net
    from oracle.make_golden import sd_digest
    z = np.load(os.path.join(golden_dir, "model_%s_%d.npz" % (name, size)))
    cfg = get_cfg(input_size=size)
    arch = get_arch(name)
    torch.manual_seed(0)
    model = get_pose_net(cfg, False, arch)
    synth.randomize_bn_(model, 1)
    sd = model.state_dict()
    assert sd_digest(sd) == str(z["digest"]), "drop-in module must create identical seeded weights"
    x = synth.make_frames(1, size, seed=11)
    with torch.no_grad():
        outs = model_ref.forward(sd, arch, x)
    for a, k in zip(outs, ("out0", "out1")):
        assert np.allclose(a.numpy(), z[k], atol=1e-5)
        # the stated tolerance must hold for the reference's own fp16-eager path
        tol = 2e-3 * np.abs(z[k]).max() + 1e-4
        assert np.abs(z[k + "_fp16eager"] - z[k]).max() <= tol


@pytest.mark.parametrize("flip,proj", [(1, 1), (0, 1), (1, 0)])
def test_glue(golden_dir, flip, proj):
    z = np.load(os.path.join(golden_dir, "glue_flip%d_proj%d.npz" % (flip, proj)))
    cfg = get_cfg(input_size=64, flip_test=bool(flip), project2image=bool(proj))
    calls = []

    def fake(img):
        calls.append(1)
        k = ("a0", "a1") if len(calls) == 1 else ("b0", "b1")
        return [torch.from_numpy(z[k[0]]), torch.from_numpy(z[k[1]])]

    _, h, t = glue_ref.multi_stage_outputs(cfg, fake, torch.zeros(2, 3, 64, 64), bool(flip), bool(proj), (64, 64))
    if not proj:
        t = [glue_ref.bilinear(x, h[0].shape[2:]) for x in t]
    fh, tg = glue_ref.aggregate(cfg, h, t)
    assert np.allclose(fh.numpy(), z["final_heatmaps"], atol=2e-6)
    assert np.allclose(tg.numpy(), z["tags"], atol=2e-6)


@pytest.mark.parametrize("case", PARSER_CASES, ids=[c[0] for c in PARSER_CASES])
def test_parser(golden_dir, case):
    name, ds, h, w, t, people, seed = case
    z = np.load(os.path.join(golden_dir, "parser_%s.npz" % name))
    cfg = get_cfg(dataset=ds, input_size=256)
    nj = cfg.DATASET.NUM_JOINTS
    det, tag = synth.plant_crowd(nj, h, w, t, num_people=people, seed=seed)
    assert hashlib.sha256(det.tobytes() + tag.tobytes()).hexdigest() == str(z["in_digest"])
    p = group_ref.HeatmapParser(cfg)
    top = p.top_k(det[None], tag[None])
    assert_topk_equal(top, {k: z[k] for k in ("val_k", "loc_k", "tag_k")}, name)
    for adj, ref in ((True, True), (True, False), (False, False)):
        ans, scores = p.parse(det[None].copy(), tag[None].copy(), adj, ref)
        a = np.array(ans[0], dtype=np.float32).reshape(-1, nj, 3 + t)
        assert np.array_equal(a, z["ans_a%d_r%d" % (adj, ref)])
        assert np.array_equal(np.array(scores, np.float32), z["scores_a%d_r%d" % (adj, ref)])


def test_golden_present(golden_dir):
    assert len(glob.glob(os.path.join(golden_dir, "*.npz"))) >= 15


def test_glue_multiscale(golden_dir):
    """oracle/glue_ref.multi_scale (valid.py:205-225 + inference.py:176-208) against the outputs of the reference's own loop"""
    from oracle.make_golden import MULTISCALE_CASES, FakeScaleModel, multiscale_inputs
    for name, scales, proj, flip, size, seed in MULTISCALE_CASES:
        z = np.load(os.path.join(golden_dir, "glue_%s.npz" % name))
        cfg = get_cfg(input_size=size, flip_test=flip, project2image=proj)
        cfg.TEST.SCALE_FACTOR = list(scales)
        base, images = multiscale_inputs(cfg, size)
        fake = FakeScaleModel(cfg.DATASET.NUM_JOINTS, seed)
        fh, tg = glue_ref.multi_scale(cfg, fake, images, base)
        dig = hashlib.sha256(b"".join(o.numpy().tobytes() for outs in fake.log for o in outs)).hexdigest()
        assert dig == str(z["in_digest"])
        assert np.allclose(fh.numpy(), z["final_heatmaps"], atol=2e-6), name
        assert np.allclose(tg.numpy(), z["tags"], atol=2e-6), name


def test_parser_shared_tag(golden_dir):
    """MODEL.TAG_PER_JOINT=False: oracle top_k / parse (no refine) against the reference's outputs"""
    from oracle.make_golden import shared_tag_case
    z = np.load(os.path.join(golden_dir, "parser_shared_tag_p6.npz"))
    cfg = get_cfg(input_size=256)
    cfg.MODEL.TAG_PER_JOINT = False
    det, tag = shared_tag_case(14, 128, 160, 2, 6, 31)
    assert hashlib.sha256(det.tobytes() + tag.tobytes()).hexdigest() == str(z["in_digest"])
    p = group_ref.HeatmapParser(cfg)
    assert_topk_equal(p.top_k(det[None], tag[None]), {k: z[k] for k in ("val_k", "loc_k", "tag_k")}, "shared tag")
    for adj in (True, False):
        ans, scores = p.parse(det[None].copy(), tag[None].copy(), adj, False)
        assert np.array_equal(np.array(ans[0], dtype=np.float32).reshape(-1, 14, 5), z["ans_a%d_r0" % adj])
        assert np.array_equal(np.array(scores, np.float32), z["scores_a%d_r0" % adj])
    ans, _ = p.parse(det[None].copy(), tag[None].copy(), True, True)      # the tiled-tag refine of the oracle runs
    assert np.array(ans[0]).shape[1:] == (14, 5)


def _glue_cfg_inputs(z, case):
    from oracle.make_golden import GLUE_CFG_SIZE, FakeScaleModel, glue_cfg
    name, center, ignore, per_joint, proj, seed = case
    cfg = glue_cfg(center, ignore, per_joint, proj)
    fake = FakeScaleModel(cfg.DATASET.NUM_JOINTS, seed, None if per_joint else 1)
    img = torch.zeros(2, 3, GLUE_CFG_SIZE, GLUE_CFG_SIZE)
    return cfg, fake, img


def test_glue_cfg_branches(golden_dir):
    """WITH_CENTER (kept / ignored) and TAG_PER_JOINT off: oracle glue against the unmodified reference's outputs"""
    from oracle.make_golden import GLUE_CFG_CASES, GLUE_CFG_SIZE
    for case in GLUE_CFG_CASES:
        z = np.load(os.path.join(golden_dir, "glue_cfg_%s.npz" % case[0]))
        cfg, fake, img = _glue_cfg_inputs(z, case)
        _, h, t = glue_ref.multi_stage_outputs(cfg, fake, img, True, case[4], (GLUE_CFG_SIZE, GLUE_CFG_SIZE))
        fh, tg = glue_ref.aggregate(cfg, h, t)
        dig = hashlib.sha256(b"".join(o.numpy().tobytes() for outs in fake.log for o in outs)).hexdigest()
        assert dig == str(z["in_digest"])
        assert np.allclose(fh.numpy(), z["final_heatmaps"], atol=2e-6), case[0]
        assert np.allclose(tg.numpy(), z["tags"], atol=2e-6), case[0]
