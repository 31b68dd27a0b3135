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


CHECK = r'''
import sys, types
import _init_paths                                   # what valid.py:24 does: <reference>/lib to the FRONT of sys.path
assert sys.path[0].endswith("lib"), sys.path[:3]
from litepose_b200.config import FLIP_CONFIG
ds = types.ModuleType("dataset"); dst = types.ModuleType("dataset.transforms"); dst.FLIP_CONFIG = FLIP_CONFIG
ds.transforms = dst; sys.modules["dataset"] = ds; sys.modules["dataset.transforms"] = dst
import models
from core.group import HeatmapParser
from core.inference import get_multi_stage_outputs
from utils.transforms import get_final_preds
import utils.zipreader
repo = "%(root)s"
assert models.pose_mobilenet.__file__.startswith(repo), models.pose_mobilenet.__file__
assert sys.modules["core.group"].__file__.startswith(repo)
assert sys.modules["utils.transforms"].__file__.startswith(repo)
assert sys.modules["core.inference"].__file__.startswith("%(ref)s")
assert utils.zipreader.__file__.startswith("%(ref)s")
import models.pose_higher_hrnet as hr                # the rest of the model zoo still comes from the reference
assert hr.__file__.startswith("%(ref)s")
print("init_paths dropin ok")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lib", "core")), reason="reference tree not present")
@pytest.mark.parametrize("how", ["sitecustomize", "runner"])
def test_dropin_survives_init_paths(tmp_path, how):
    """valid.py's `import _init_paths` inserts <reference>/lib at sys.path[0] AFTER PYTHONPATH /
    sitecustomize, and <reference>/lib/models is a regular package - plain path shadowing loses and the model would
    silently be the stock eager one.  The meta-path drop-in (litepose_b200/dropin.py) must win in both launch modes."""
    script = tmp_path / "check_like_valid.py"
    script.write_text(CHECK % {"ref": REF, "root": ROOT})
    env = dict(os.environ)
    if how == "sitecustomize":
        env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "litepose_b200", "dropin_site"), ROOT])
        # cwd = reference root so that `import _init_paths` resolves like for `python valid.py`
        cmd = [sys.executable, "-c", "import sys; sys.path.insert(0, %r); exec(open(%r).read())" % (REF, str(script))]
    else:
        env["PYTHONPATH"] = ROOT
        # the runner puts the script's directory first (as `python script.py` does); _init_paths lives beside valid.py
        link = tmp_path / "ref"
        os.symlink(REF, link)
        runner_script = tmp_path / "run_in_ref.py"
        runner_script.write_text("import sys; sys.path.insert(0, %r)\n" % REF + CHECK % {"ref": REF, "root": ROOT})
        cmd = [sys.executable, "-m", "litepose_b200.dropin", str(runner_script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=REF)
    assert r.returncode == 0 and "init_paths dropin ok" in r.stdout, r.stdout + r.stderr
