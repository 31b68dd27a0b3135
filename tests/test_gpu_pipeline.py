"""-m gpu: the public end-to-end call (LitePosePipeline.step: pinned host frames -> keypoints on the host) on the bench
workload, against the oracle pipeline the bench times as its CPU arm (bench.py:cpu_reference_step = reference
valid.py:195-229: two forwards with flip, glue, parser per image)."""
import numpy as np
import pytest
import torch

from litepose_b200 import _lib, synth
from litepose_b200.config import get_arch, get_cfg
from litepose_b200.lib.models.pose_mobilenet import get_pose_net
from litepose_b200.pipeline import LitePosePipeline, PlantedCrowd
from oracle import group_ref

pytestmark = pytest.mark.gpu


def _setup(arch_name, size, n, people, keep=64, seed=77):
    cfg = get_cfg(input_size=size)
    arch = get_arch(arch_name)
    torch.manual_seed(0)
    model = synth.scale_heads_(synth.randomize_bn_(get_pose_net(cfg, False, arch), 1)).eval()
    sd = {k: v.float().clone() for k, v in model.state_dict().items()}
    frames = synth.make_frames(n, size, seed=1234)
    pipe = LitePosePipeline(model.cuda(), cfg, use_graphs=True, keep=keep)
    plant_dev = PlantedCrowd(n, 14, size, size, 2, num_people=people, seed=seed, device="cuda")
    plant_cpu = PlantedCrowd(n, 14, size, size, 2, num_people=people, seed=seed, device="cpu")
    return cfg, arch, sd, frames, pipe, plant_dev, plant_cpu


def test_step_vs_oracle_pipeline_s512():
    """BASELINE config 3 geometry (LitePose-S 512^2, flip test, PROJECT2IMAGE, adjust + refine), batch 2."""
    import bench
    n, size = 2, 512
    cfg, arch, sd, frames, pipe, plant_dev, plant_cpu = _setup("S", size, n, 5)
    got = pipe.step(frames.half().pin_memory(), plant_dev)
    got2 = pipe.step(frames.half().pin_memory(), plant_dev)            # graph replay
    st = pipe._get_state(n, size, size, torch.float16, plant_dev)
    det_g, tag_g = st["det"].cpu().numpy(), st["tag"].cpu().numpy()

    # (1) the maps the device parser saw == glue(oracle forward) + plant within the fp tolerance of the model
    from oracle import glue_ref, model_ref
    with torch.no_grad():
        _, hm, tg = glue_ref.multi_stage_outputs(cfg, lambda im: model_ref.forward(sd, arch, im), frames, True, True,
                                                 (size, size))
        det_o, tag_o = glue_ref.aggregate(cfg, hm, tg)
        det_o, tag_o = det_o.contiguous(), tag_o.contiguous()
        det_raw, tag_raw = det_o.clone(), tag_o.clone()
        plant_cpu.apply(det_o, tag_o)
    for name, g, o, raw in (("det", det_g, det_o.numpy(), det_raw.numpy()), ("tag", tag_g, tag_o.numpy(), tag_raw.numpy())):
        err = np.abs(g - o).max()
        lim = 2e-3 * np.abs(raw).max() + 1e-4          # tolerance relative to the NETWORK's maps, not the planted peaks
        assert err <= lim, "%s: %.3e > %.3e" % (name, err, lim)

    # (2) the packed result of the step is bit-exactly the oracle parser's answer on the maps the device produced
    op = group_ref.HeatmapParser(cfg)
    for i in range(n):
        ans, scores = op.parse(det_g[i:i + 1].copy(), tag_g[i:i + 1].copy(), True, True)
        e = np.asarray(ans[0], np.float32).reshape(-1, 14, 5)
        for res in (got, got2):
            assert res[i][2] == e.shape[0]
            assert np.array_equal(res[i][0], e), i
            assert np.array_equal(np.asarray(res[i][1], np.float32), np.asarray(scores, np.float32))

    # (3) against the full CPU oracle pipeline (fp32 forward): same person count, same integer peak locations and
    # adjusted coordinates for every joint the matcher assigned (tag columns non-zero), values / tags within tolerance.
    # Joints filled in by refine on planes where the person has no planted peak are an argmax over network noise and
    # are compared through (2) only.
    exp = bench.cpu_reference_step(cfg, arch, sd, frames, plant_cpu)
    for i in range(n):
        e = np.asarray(exp[i][0][0], np.float32).reshape(-1, 14, 5)
        a = got[i][0]
        assert a.shape == e.shape, (i, a.shape, e.shape)
        matched = (e[:, :, 3] != 0) | (e[:, :, 4] != 0)
        assert np.array_equal(matched, (a[:, :, 3] != 0) | (a[:, :, 4] != 0))
        assert matched.sum() >= 5 * 10
        assert np.array_equal(a[matched][:, :2], e[matched][:, :2]), "peak locations"
        assert np.abs(a[matched][:, 2:] - e[matched][:, 2:]).max() <= 2e-3 * np.abs(e[matched][:, 2:]).max() + 1e-4


def test_crowd_overflow_returns_every_person():
    """BASELINE config 5 geometry (LitePose-L 640^2, 30-person planted crowd): more persons than the 64-person packed
    payload -> the step still returns all of them (reference lib/core/group.py:96 keeps every person), equal to the
    oracle parser on the same maps; unpack() without the second-chance copy raises instead of clipping."""
    n, size = 2, 640
    cfg, arch, sd, frames, pipe, plant_dev, _ = _setup("L", size, n, 30, keep=24, seed=78)
    got = pipe.step(frames.half().pin_memory(), plant_dev)
    st = pipe._get_state(n, size, size, torch.float16, plant_dev)
    det_g, tag_g = st["det"].cpu().numpy(), st["tag"].cpu().numpy()
    op = group_ref.HeatmapParser(cfg)
    counts = []
    for i in range(n):
        ans, scores = op.parse(det_g[i:i + 1].copy(), tag_g[i:i + 1].copy(), True, True)
        e = np.asarray(ans[0], np.float32).reshape(-1, 14, 5)
        counts.append(e.shape[0])
        assert got[i][2] == e.shape[0] and got[i][0].shape == e.shape
        assert np.array_equal(got[i][0], e)
        assert np.array_equal(np.asarray(got[i][1], np.float32), np.asarray(scores, np.float32))
    assert max(counts) > 24, counts
    with pytest.raises(_lib.LitePoseError):
        pipe.unpack(st["host"], st["row"], st["T"])


def test_unsupported_cfg_is_rejected():
    cfg = get_cfg(input_size=128)
    torch.manual_seed(0)
    model = get_pose_net(cfg, False, get_arch("XS")).eval().cuda()
    def broken_reference_combo(c):                 # the reference drops the only tag map there (inference.py:147-150)
        c.DATASET.WITH_CENTER, c.TEST.IGNORE_CENTER, c.MODEL.TAG_PER_JOINT = True, True, False

    for mutate in (broken_reference_combo, lambda c: setattr(c.TEST, "WITH_HEATMAPS", (True, False)),
                   lambda c: setattr(c.TEST, "WITH_AE", (True, True))):
        c = get_cfg(input_size=128)
        mutate(c)
        with pytest.raises(NotImplementedError):
            LitePosePipeline(model, c)
    c = get_cfg(input_size=128)
    c.TEST.SCALE_FACTOR = [0.5, 2]            # the reference needs the scale-1 pass for the tags (valid.py:224)
    with pytest.raises(ValueError):
        LitePosePipeline(model, c)
    c.TEST.SCALE_FACTOR = [0.5, 1]
    pipe = LitePosePipeline(model, c)
    with pytest.raises(RuntimeError):         # the single-scale entry points refuse a multi-scale cfg
        pipe.step(synth.make_frames(1, 128, seed=1).half().pin_memory())


def test_plant_change_recaptures_graph():
    """a captured step graph must not keep replaying the first call's plant hook."""
    n, size = 2, 128
    cfg, arch, sd, frames, pipe, plant_a, _ = _setup("XS", size, n, 3, seed=5)
    plant_b = PlantedCrowd(n, 14, size, size, 2, num_people=2, seed=6, device="cuda")
    fr = frames.half().pin_memory()
    a1 = pipe.step(fr, plant_a)
    b = pipe.step(fr, plant_b)
    none = pipe.step(fr, None)
    a2 = pipe.step(fr, plant_a)
    assert [r[2] for r in none] == [0, 0]
    assert [r[2] for r in a1] == [r[2] for r in a2] and all(np.array_equal(x[0], y[0]) for x, y in zip(a1, a2))
    assert [r[2] for r in b] != [r[2] for r in a1] or not np.array_equal(a1[0][0], b[0][0])


def test_submit_collect_equals_step():
    """The asynchronous API (two steps in flight, result slots, no per-step synchronisation) returns exactly what the
    blocking step() returns, in submission order, also when the inputs alternate."""
    n, size = 2, 128
    cfg, arch, sd, frames, pipe, plant, _ = _setup("XS", size, n, 3, seed=5)
    fa = frames.half().pin_memory()
    fb = synth.make_frames(n, size, seed=99).half().pin_memory()
    ref_a, ref_b = pipe.step(fa, plant), pipe.step(fb, plant)
    seq = [fa, fb, fb, fa, fa, fb]
    exp = [ref_a, ref_b, ref_b, ref_a, ref_a, ref_b]
    got, prev = [], None
    for f in seq:
        t = pipe.submit(f, plant)
        if prev is not None:
            got.append(pipe.collect(prev)[0])
        prev = t
    got.append(pipe.collect(prev)[0])
    with pytest.raises(RuntimeError):
        pipe.collect(prev)
    for g, e in zip(got, exp):
        assert [r[2] for r in g] == [r[2] for r in e]
        for x, y in zip(g, e):
            assert np.array_equal(x[0], y[0]) and np.array_equal(np.asarray(x[1]), np.asarray(y[1]))


@pytest.mark.parametrize("scales,proj", [([0.5, 1, 2], True), ([1, 2], False)])
def test_multiscale_step_vs_oracle(scales, proj):
    """TEST.SCALE_FACTOR with several entries (reference valid.py:198-229): per scale two network passes + one
    accumulating glue launch, then the parser.  Maps within the model tolerance of the fp32 oracle loop
    (oracle/glue_ref.multi_scale on oracle/model_ref.forward); keypoints bit-exactly the oracle parser's answer on the
    maps the device produced."""
    from oracle import glue_ref, model_ref
    n, size = 2, 128
    cfg = get_cfg(input_size=size, project2image=proj)
    cfg.TEST.SCALE_FACTOR = list(scales)
    arch = get_arch("XS")
    torch.manual_seed(0)
    model = synth.scale_heads_(synth.randomize_bn_(get_pose_net(cfg, False, arch), 1)).eval()
    sd = {k: v.float().clone() for k, v in model.state_dict().items()}
    frames = {float(s): synth.make_frames(n, int(size * s), seed=40 + i) for i, s in enumerate(scales)}
    pipe = LitePosePipeline(model.cuda(), cfg, use_graphs=True)
    big = int(size * max(scales))
    Hd = size if proj else big // 2
    plant_dev = PlantedCrowd(n, 14, Hd, Hd, 2, num_people=3, seed=9, device="cuda")
    plant_cpu = PlantedCrowd(n, 14, Hd, Hd, 2, num_people=3, seed=9, device="cpu")
    got = pipe.step_multiscale({s: f.half().pin_memory() for s, f in frames.items()}, plant_dev)
    st = pipe._get_state(n, size, size, torch.float16, plant_dev, det_hw=(Hd, Hd))
    det_g, tag_g = st["det"].cpu().numpy(), st["tag"].cpu().numpy()
    with torch.no_grad():
        det_o, tag_o = glue_ref.multi_scale(cfg, lambda im: model_ref.forward(sd, arch, im), frames, (size, size))
        det_o, tag_o = det_o.contiguous(), tag_o.contiguous()
        det_raw, tag_raw = det_o.clone(), tag_o.clone()
        plant_cpu.apply(det_o, tag_o)
    assert det_g.shape == tuple(det_o.shape) and tag_g.shape == tuple(tag_o.shape)
    for name, g, o, raw in (("det", det_g, det_o.numpy(), det_raw.numpy()), ("tag", tag_g, tag_o.numpy(), tag_raw.numpy())):
        err = np.abs(g - o).max()
        lim = 2e-3 * np.abs(raw).max() + 1e-4
        assert err <= lim, "%s: %.3e > %.3e" % (name, err, lim)
    op = group_ref.HeatmapParser(cfg)
    for i in range(n):
        ans, scores = op.parse(det_g[i:i + 1].copy(), tag_g[i:i + 1].copy(), True, True)
        e = np.asarray(ans[0], np.float32).reshape(-1, 14, 5)
        assert got[i][2] == e.shape[0] and e.shape[0] >= 3
        assert np.array_equal(got[i][0], e), i
        assert np.array_equal(np.asarray(got[i][1], np.float32), np.asarray(scores, np.float32))


@pytest.mark.parametrize("center,ignore,per_joint,dataset", [(True, True, True, "crowd_pose"), (True, False, True, "coco"),
                                                             (False, True, False, "crowd_pose")])
def test_step_cfg_branches_vs_oracle(center, ignore, per_joint, dataset):
    """DATASET.WITH_CENTER (centre joint ignored / kept) and MODEL.TAG_PER_JOINT off, end to end: the network with the
    matching head widths, the general glue entry and the parser, against the oracle (maps within the model tolerance,
    keypoints bit-exactly the oracle parser's answer on the device's maps).  Heat peaks are planted, tags are the
    network's own.  The kept centre joint is a COCO case: the reference's joint_order puts it at index 17
    (group.py:113-118), which a 15-joint CrowdPose model does not have."""
    from oracle import glue_ref, model_ref
    from oracle.make_golden import glue_cfg
    n, size = 2, 128
    cfg = glue_cfg(center, ignore, per_joint, True, size=size, dataset=dataset)
    arch = get_arch("XS")
    torch.manual_seed(0)
    model = synth.scale_heads_(synth.randomize_bn_(get_pose_net(cfg, False, arch), 1)).eval()
    sd = {k: v.float().clone() for k, v in model.state_dict().items()}
    frames = synth.make_frames(n, size, seed=21)
    pipe = LitePosePipeline(model.cuda(), cfg, use_graphs=True)
    J = pipe.params.num_joints
    assert J == {"crowd_pose": 14, "coco": 17}[dataset] + (1 if (center and not ignore) else 0)
    plants = []
    for dev in ("cuda", "cpu"):
        pl = PlantedCrowd(n, J, size, size, 2, num_people=3, seed=12, device=dev)
        pl.tidx, pl.tval = pl.tidx[:0], pl.tval[:0]          # heat peaks only
        plants.append(pl)
    got = pipe.step(frames.half().pin_memory(), plants[0])
    st = pipe._get_state(n, size, size, torch.float16, plants[0])
    det_g, tag_g = st["det"].cpu().numpy(), st["tag"].cpu().numpy()
    with torch.no_grad():
        _, hm, tg = glue_ref.multi_stage_outputs(cfg, lambda im: model_ref.forward(sd, arch, im), frames, True, True,
                                                 (size, size))
        det_o, tag_o = glue_ref.aggregate(cfg, hm, tg)
        det_o, tag_o = det_o.contiguous(), tag_o.contiguous()
        det_raw = det_o.clone()
        plants[1].apply(det_o, tag_o)
    assert det_g.shape == tuple(det_o.shape) and tag_g.shape == tuple(tag_o.shape)
    assert tag_g.shape[1] == (J if per_joint else 1)
    assert np.abs(det_g - det_o.numpy()).max() <= 2e-3 * np.abs(det_raw.numpy()).max() + 1e-4
    assert np.abs(tag_g - tag_o.numpy()).max() <= 2e-3 * np.abs(tag_o.numpy()).max() + 1e-4
    op = group_ref.HeatmapParser(cfg)
    for i in range(n):
        ans, scores = op.parse(det_g[i:i + 1].copy(), tag_g[i:i + 1].copy(), True, True)
        e = np.asarray(ans[0], np.float32).reshape(-1, J, 5)
        assert got[i][2] == e.shape[0] and e.shape[0] >= 1
        assert np.array_equal(got[i][0], e), i
        assert np.array_equal(np.asarray(got[i][1], np.float32), np.asarray(scores, np.float32))


@pytest.mark.parametrize("scales", [[1], [0.5, 1, 2]])
def test_infer_images_is_the_valid_loop_body(scales):
    """LitePosePipeline.infer_images = valid.py:198-233 for a batch of uint8 images: per-scale warp + normalise on the
    device, network passes, glue (accumulating over the scales), parser, get_final_preds.  Against the same chain built
    from the separately pinned pieces: the device pre-processing (bit-identical to cv2 elsewhere), the oracle loop on
    those frames (maps within the model tolerance), the oracle parser on the device's maps and the host get_final_preds
    with the last scale's centre / scale (bit-exact final coordinates)."""
    from litepose_b200.lib.utils import transforms as T
    from oracle import glue_ref, model_ref
    n, H, W, size = 2, 150, 200, 128
    cfg = get_cfg(input_size=size)
    cfg.TEST.SCALE_FACTOR = list(scales)
    arch = get_arch("XS")
    torch.manual_seed(0)
    model = synth.scale_heads_(synth.randomize_bn_(get_pose_net(cfg, False, arch), 1)).eval()
    sd = {k: v.float().clone() for k, v in model.state_dict().items()}
    imgs = torch.from_numpy(np.random.RandomState(3).randint(0, 256, (n, H, W, 3)).astype(np.uint8))
    smin = min(scales)
    (bw, bh), _, _ = T.get_multi_scale_size(np.empty((H, W, 3), np.uint8), size, 1.0, smin)
    pipe = LitePosePipeline(model.cuda(), cfg, use_graphs=True)
    plant_dev = PlantedCrowd(n, 14, bh, bw, 2, num_people=3, seed=15, device="cuda")
    plant_cpu = PlantedCrowd(n, 14, bh, bw, 2, num_people=3, seed=15, device="cpu")
    got = pipe.infer_images(imgs.pin_memory(), plant=plant_dev)
    got2 = pipe.infer_images(imgs.pin_memory(), plant=plant_dev)
    # the pieces
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    frames, center, scale = {}, None, None
    for s in sorted(scales, reverse=True):
        x, center, scale = T.resize_align_normalize_device(imgs, size, s, smin, mean, std, half=True)
        frames[float(s)] = x.float().cpu()
    assert tuple(frames[1.0].shape[2:]) == (bh, bw)
    st = pipe._last_state
    det_g, tag_g = st["det"].cpu().numpy(), st["tag"].cpu().numpy()
    with torch.no_grad():
        fwd = lambda im: model_ref.forward(sd, arch, im)
        if len(scales) > 1:
            det_o, tag_o = glue_ref.multi_scale(cfg, fwd, frames, (bw, bh))
        else:
            _, hm, tg = glue_ref.multi_stage_outputs(cfg, fwd, frames[1.0], True, True, (bw, bh))
            det_o, tag_o = glue_ref.aggregate(cfg, hm, tg)
        det_o, tag_o = det_o.contiguous(), tag_o.contiguous()
        lim_d, lim_t = 2e-3 * det_o.abs().max().item() + 1e-4, 2e-3 * tag_o.abs().max().item() + 1e-4
        plant_cpu.apply(det_o, tag_o)
    assert np.abs(det_g - det_o.numpy()).max() <= lim_d and np.abs(tag_g - tag_o.numpy()).max() <= lim_t
    op = group_ref.HeatmapParser(cfg)
    for i in range(n):
        ans, scores = op.parse(det_g[i:i + 1].copy(), tag_g[i:i + 1].copy(), True, True)
        exp = T.get_final_preds(ans, center, scale, [bw, bh])
        for res in (got, got2):
            assert res[i][2] == len(exp) and len(exp) >= 3
            assert np.array_equal(res[i][0], np.stack(exp).astype(np.float32)), i
            assert np.array_equal(np.asarray(res[i][1], np.float32), np.asarray(scores, np.float32))
