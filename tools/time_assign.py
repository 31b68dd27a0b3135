"""Times the fast_utils parser (find_peaks + KM assign) on the projected maps of the bench workload."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from litepose_b200 import synth
from litepose_b200.config import get_cfg
from litepose_b200.fast_utils.group import HeatmapParser as FastParser
B, S = 32, 512
cfg = get_cfg(input_size=S, flip_test=False, adjust=False, refine=False)
det, tag = synth.plant_crowd_batch(B, 14, S, S, 1, num_people=5, seed=77)
det, tag = torch.from_numpy(det).cuda(), torch.from_numpy(tag).cuda()
fp = FastParser(cfg)
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t = timed(lambda: fp.parse_batch(det, tag))
num, _ = fp.parse_batch(det, tag)
print(json.dumps({"fast_utils_parse_ms_per_32_frames": t, "persons": num[:8].tolist()}))
