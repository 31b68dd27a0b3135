"""sub-network extraction from a larger checkpoint (reference weight_transfer.py:75-146)."""
import ast
import os

import pytest
import torch

from litepose_b200.config import get_arch, get_cfg
from litepose_b200.lib.models.pose_mobilenet import get_pose_net
from litepose_b200.subnet import extract_subnet, extract_subnet_state_dict

REF = "/root/reference/weight_transfer.py"


def _big(seed=0):
    torch.manual_seed(seed)
    net = get_pose_net(get_cfg(), False, get_arch("L"))
    for m in net.modules():                                      # non-trivial BN statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.75, 1.25)
    return net.eval()


@pytest.mark.parametrize("arch", ["XS", "S", "M"])
def test_prefix_slices_and_strict_load(arch):
    big = _big()
    sub = extract_subnet(big.state_dict(), get_cfg(), get_arch(arch))
    sd, sup = sub.state_dict(), big.state_dict()
    assert len(sd) == 679
    for k, v in sd.items():
        assert torch.equal(v, sup[k][tuple(slice(0, d) for d in v.shape)].to(v.dtype)), k
    # a '1.'-prefixed (network_to_half) checkpoint works too; too small a source is refused
    pref = {"1." + k: v for k, v in sup.items()}
    again = extract_subnet_state_dict(pref, sd)
    assert all(torch.equal(again[k], sd[k]) for k in sd)
    with pytest.raises(ValueError):
        extract_subnet_state_dict(sd, sup)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present")
def test_matches_reference_transfer():
    """the reference's own transfer() (function bodies executed from weight_transfer.py where it lies) on reference modules"""
    from oracle import refshim
    src = open(REF).read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name.startswith("transfer")]
    cfg = get_cfg()
    ns = {"cfg": cfg}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "weight_transfer_fns", "exec"), ns)
    big = refshim.build_reference_model(cfg, get_arch("L"), seed=3)
    tiny = {"img_size": 256, "input_channel": 16, "deconv_setting": [16, 24, 24],
            "backbone_setting": [{"num_blocks": 2, "stride": 2, "channel": 16, "block_setting": [[6, 7]] * 2},
                                 {"num_blocks": 3, "stride": 2, "channel": 24, "block_setting": [[6, 7]] * 3},
                                 {"num_blocks": 2, "stride": 2, "channel": 40, "block_setting": [[6, 7]] * 2},
                                 {"num_blocks": 1, "stride": 1, "channel": 64, "block_setting": [[6, 7]] * 1}]}
    for arch in (get_arch("S"), tiny):
        small = refshim.build_reference_model(cfg, arch, seed=4)
        ns["transfer"](big, small, arch)
        want = small.state_dict()
        got = extract_subnet_state_dict(big.state_dict(), want)
        for k in want:
            if k.endswith("num_batches_tracked"):
                continue                                   # the reference leaves the target's own counter
            assert torch.equal(got[k], want[k]), k
