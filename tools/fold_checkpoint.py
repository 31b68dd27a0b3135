"""Offline conversion of a LitePose checkpoint into the folded, kernel-packed format:

    python tools/fold_checkpoint.py --arch S --state-dict model.pth --out litepose_s.folded.npz
    python tools/fold_checkpoint.py --arch XS --random --out xs_random.folded.npz      # seeded random-init weights

BN folding (reference fuse_bn.py:81-162) and weight packing run once here (CPU is enough); at load time
``LitePoseEngine.from_folded(path, "cuda")`` only uploads the arrays."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from litepose_b200.config import get_arch, get_cfg  # noqa: E402
from litepose_b200.engine import LitePoseEngine  # noqa: E402
from litepose_b200.lib.models.pose_mobilenet import get_pose_net  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="S", help="XS|S|M|L or a path to an arch json")
    ap.add_argument("--dataset", default="crowd_pose")
    ap.add_argument("--state-dict", help=".pth with the reference's state_dict (keys optionally prefixed '1.' by network_to_half)")
    ap.add_argument("--supernet", help=".pth of a larger (super)network: the --arch sub-network is extracted by prefix slicing "
                                       "(reference weight_transfer.py:75-146)")
    ap.add_argument("--random", action="store_true", help="seeded random-init weights (synthetic benchmarks)")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    arch = get_arch(a.arch)
    cfg = get_cfg(dataset=a.dataset)
    torch.manual_seed(0)
    model = get_pose_net(cfg, False, arch).eval()
    if a.state_dict:
        sd = torch.load(a.state_dict, map_location="cpu")
        sd = {(k[2:] if k.startswith("1.") else k): v for k, v in sd.items()}      # weight_transfer.py:199-200
        model.load_state_dict(sd, strict=True)
    elif a.supernet:
        from litepose_b200.subnet import extract_subnet_state_dict
        sup = torch.load(a.supernet, map_location="cpu")
        model.load_state_dict(extract_subnet_state_dict(sup, model.state_dict()), strict=True)
    elif not a.random:
        ap.error("give --state-dict, --supernet or --random")
    eng = LitePoseEngine(model.state_dict(), arch, "cpu")
    eng.export_folded(a.out)
    print(a.out, os.path.getsize(a.out), "bytes")


if __name__ == "__main__":
    main()
