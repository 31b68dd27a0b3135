"""CPU, build container only (skipped where /root/reference is absent): the sys.path shadowing
of INTEGRATION.md §1 resolves `models.pose_mobilenet` / `core.group` / `utils.transforms` to this repo and
`core.inference`, `utils.zipreader` ... to the unmodified reference, and the reference glue runs on our module's outputs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = r'''
import sys, types
sys.path.insert(0, "%(ref)s/lib"); sys.path.insert(0, "%(ref)s")
sys.path.insert(0, "%(root)s"); sys.path.insert(0, "%(root)s/litepose_b200/lib")
from litepose_b200.config import FLIP_CONFIG, get_cfg, get_arch
ds = types.ModuleType("dataset"); dst = types.ModuleType("dataset.transforms"); dst.FLIP_CONFIG = FLIP_CONFIG
ds.transforms = dst; sys.modules["dataset"] = ds; sys.modules["dataset.transforms"] = dst
import torch, models
from core.group import HeatmapParser
from core.inference import get_multi_stage_outputs, aggregate_results
assert "litepose_b200" in models.pose_mobilenet.__file__, models.pose_mobilenet.__file__
assert "litepose_b200" in sys.modules["core.group"].__file__
assert "%(ref)s" in sys.modules["core.inference"].__file__
cfg = get_cfg(input_size=64)
model = eval("models." + cfg.MODEL.NAME + ".get_pose_net")(cfg, is_train=True, cfg_arch=get_arch("XS")).eval()
with torch.no_grad():
    o, h, t = get_multi_stage_outputs(cfg, model, torch.rand(1, 3, 64, 64), True, True, (64, 64))
    fh, tl = aggregate_results(cfg, 1, None, [], h, t)
assert tuple(fh.shape) == (1, 14, 64, 64) and tuple(torch.cat(tl, 4).shape) == (1, 14, 64, 64, 2)
HeatmapParser(cfg)
# valid.py:44-46 and :27-29: the coordinate helpers come from this repo, the other utils modules from the reference
from utils.transforms import resize_align_multi_scale, get_final_preds, get_multi_scale_size
import utils.transforms, utils.zipreader
assert "litepose_b200" in utils.transforms.__file__ and "%(ref)s" in utils.zipreader.__file__
assert utils.transforms.fliplr_joints.__module__ == "_reference_utils_transforms"      # delegated, not restated
import numpy as np
size, center, scale = get_multi_scale_size(np.zeros((480, 640, 3), np.uint8), 512, 1.0, 1.0)
assert size == (704, 512) and list(center) == [320, 240]
print("dropin ok")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lib", "core")), reason="reference tree not present")
def test_sys_path_shadowing():
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"ref": REF, "root": ROOT}], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "dropin ok" in r.stdout, r.stdout + r.stderr
