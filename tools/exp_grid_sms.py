"""Experiment: persistent kernels sized for a fraction of the SMs (LP_GRID_SMS) so that the two passes of the flip test,
which run on two streams, occupy disjoint SMs instead of alternating whole-chip kernels.  One JSON line per setting."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from litepose_b200 import synth  # noqa: E402
from litepose_b200.config import get_arch, get_cfg  # noqa: E402
from litepose_b200.lib.models.pose_mobilenet import get_pose_net  # noqa: E402
from litepose_b200.pipeline import LitePosePipeline, PlantedCrowd  # noqa: E402

dev = torch.device("cuda", 0)
cfg = get_cfg(input_size=512)
torch.manual_seed(0)
model = synth.scale_heads_(synth.randomize_bn_(get_pose_net(cfg, False, get_arch("S")), 1)).eval().to(dev)
x = synth.make_frames(32, 512, seed=1234).half().to(dev)
plant = PlantedCrowd(32, 14, 512, 512, 2, num_people=5, seed=77, device=dev)
ref = None
for sms in [int(v) for v in (sys.argv[1:] or ["148", "74", "111", "96", "128", "148"])]:
    os.environ["LP_GRID_SMS"] = str(sms)
    pipe = LitePosePipeline(model, cfg, use_graphs=True)
    for _ in range(5):
        packed, ev = pipe.step_device_overlapped(x, plant)
    torch.cuda.synchronize()
    if ref is None:
        ref = packed.clone()
    same = bool(torch.equal(ref, packed))
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            _, ev = pipe.step_device_overlapped(x, plant)
        torch.cuda.current_stream().wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        best = ms if best is None else min(best, ms)
    print(json.dumps({"grid_sms": sms, "ms_per_step": best, "frames_per_s": 32e3 / best, "payload_identical": same}), flush=True)
    del pipe
