"""Runs only bench.py's secondary variants (nano-demo settings, fast_utils parser, uint8 loop, multi-scale test) on one
GPU and prints them as JSON - a development helper; the numbers of record are the ones inside the bench line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from litepose_b200 import synth  # noqa: E402
from litepose_b200.config import get_arch, get_cfg  # noqa: E402
from litepose_b200.lib.models.pose_mobilenet import get_pose_net  # noqa: E402
from litepose_b200.pipeline import PlantedCrowd  # noqa: E402

args = argparse.Namespace(size=512, people=5, batch=32, arch="S")
dev = torch.device("cuda", 0)
cfg = get_cfg(input_size=args.size)
torch.manual_seed(0)
model = synth.scale_heads_(synth.randomize_bn_(get_pose_net(cfg, False, get_arch(args.arch)), 1)).eval().to(dev)
x = synth.make_frames(args.batch, args.size, seed=1234).half().to(dev)
plant = PlantedCrowd(args.batch, 14, args.size, args.size, 2, num_people=args.people, seed=77, device=dev)
print(json.dumps(bench.time_variants(args, model, x, plant, dev)))
