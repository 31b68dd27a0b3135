"""Small launches of the round-2 kernels (block_s1 resident + streaming, fused stem, deferred-epilogue head, pipeline step)
for compute-sanitizer:  compute-sanitizer --tool memcheck python tools/memcheck_r2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from litepose_b200 import _lib, synth
from litepose_b200.config import get_arch, get_cfg
from litepose_b200.lib.models.pose_mobilenet import get_pose_net
from litepose_b200.pipeline import LitePosePipeline, PlantedCrowd

lib = _lib.load()
cfg = get_cfg(input_size=128)
torch.manual_seed(0)
for arch in ("S", "M"):
    model = synth.scale_heads_(synth.randomize_bn_(get_pose_net(cfg, False, get_arch(arch)), 1)).eval().cuda()
    x = synth.make_frames(3, 128, seed=5).cuda().half()
    eng = model.lp_engine()
    for flip in (False, True):
        o = eng.run(x, flip=flip)
        assert torch.isfinite(o[0]).all() and torch.isfinite(o[1]).all()
    print(arch, "forward ok, launches", lib.lp_launch_count())
pipe = LitePosePipeline(model, cfg, use_graphs=False)
plant = PlantedCrowd(3, 14, 128, 128, 2, num_people=3, seed=4, device="cuda")
res = pipe.step(synth.make_frames(3, 128, seed=7).half().pin_memory(), plant)
print("pipeline persons", [r[2] for r in res])
torch.cuda.synchronize()
print("done")
