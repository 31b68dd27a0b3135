"""One benchmark step (LitePose-S 512x512, batch 32, flip + glue + parser) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off` (launch list / full captures).
Not a timing tool: numbers printed under a profiler are never bench values."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from litepose_b200 import synth  # noqa: E402
from litepose_b200.config import get_arch, get_cfg  # noqa: E402
from litepose_b200.lib.models.pose_mobilenet import get_pose_net  # noqa: E402
from litepose_b200.pipeline import LitePosePipeline, PlantedCrowd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="S")
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--people", type=int, default=5)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()

dev = torch.device("cuda", 0)
cfg = get_cfg(input_size=a.size)
torch.manual_seed(0)
model = synth.scale_heads_(synth.randomize_bn_(get_pose_net(cfg, False, get_arch(a.arch)), 1)).eval().to(dev)
pipe = LitePosePipeline(model, cfg, use_graphs=False)
x = synth.make_frames(a.batch, a.size, seed=1234).half().to(dev)
plant = PlantedCrowd(a.batch, 14, a.size, a.size, 2, num_people=a.people, seed=77, device=dev)
pipe.step_device(x, plant)
pipe.step_device(x, plant)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    pipe.step_device(x, plant)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled %d step(s)" % a.steps)
