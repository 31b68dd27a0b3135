"""Small launches of the kernels added in the second session of round 2 (general / multi-scale glue, 64-wide tag matcher,
payload pack, planted crowd) for compute-sanitizer:  compute-sanitizer --tool memcheck python tools/memcheck_s2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from litepose_b200 import _lib, synth
from litepose_b200.config import get_arch, get_cfg
from litepose_b200.lib.core.group import HeatmapParser
from litepose_b200.lib.models.pose_mobilenet import get_pose_net
from litepose_b200.pipeline import LitePosePipeline, PlantedCrowd
from oracle.make_golden import glue_cfg

lib = _lib.load()
torch.manual_seed(0)
# multi-scale step (glue_kernel: accumulate / divide / no-tag modes, tile sides 64..16), pack + plant kernels
cfg = get_cfg(input_size=128)
cfg.TEST.SCALE_FACTOR = [0.5, 1, 1.5]
model = synth.scale_heads_(synth.randomize_bn_(get_pose_net(cfg, False, get_arch("XS")), 1)).eval().cuda()
pipe = LitePosePipeline(model, cfg, use_graphs=False)
plant = PlantedCrowd(2, 14, 128, 128, 2, num_people=3, seed=4, device="cuda")
frames = {s: synth.make_frames(2, int(128 * s), seed=7).half().pin_memory() for s in (0.5, 1.0, 1.5)}
print("multi-scale persons", [r[2] for r in pipe.step_multiscale(frames, plant)])
# WITH_CENTER (ignored) and shared tag map through the general glue entry
for c in ((True, True, True), (False, True, False)):
    cfg = glue_cfg(*c, True, size=128)
    model = synth.scale_heads_(synth.randomize_bn_(get_pose_net(cfg, False, get_arch("XS")), 1)).eval().cuda()
    pipe = LitePosePipeline(model, cfg, use_graphs=False)
    pl = PlantedCrowd(2, 14, 128, 128, 2, num_people=2, seed=5, device="cuda")
    pl.tidx, pl.tval = pl.tidx[:0], pl.tval[:0]
    print("cfg", c, "persons", [r[2] for r in pipe.step(synth.make_frames(2, 128, seed=8).half().pin_memory(), pl)])
# 64-wide matcher
cfg = get_cfg(input_size=192)
cfg.DATASET.MAX_NUM_PEOPLE = 48
det, tag = synth.plant_crowd_batch(2, 14, 192, 192, 2, num_people=36, seed=3)
out = HeatmapParser(cfg).parse_batch(torch.from_numpy(det).cuda(), torch.from_numpy(tag).cuda(), True, True)
print("wide matcher persons", [len(o[0][0]) for o in out])
torch.cuda.synchronize()
print("done")
