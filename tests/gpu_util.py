"""Helpers for the -m gpu parity tests: call the C ABI on torch CUDA tensors and build
fp32 CPU references (oracle side) on the same seeded inputs."""
import numpy as np
import torch
import torch.nn.functional as F

from litepose_b200 import _lib


def stream():
    return torch.cuda.current_stream().cuda_stream


def nhwc16(x_nchw):
    """fp32 NCHW CPU -> fp16 NHWC CUDA"""
    return x_nchw.permute(0, 2, 3, 1).contiguous().half().cuda()


def from_nhwc(y):
    return y.float().cpu().permute(0, 3, 1, 2).contiguous()


def q16(t):
    return t.half().float()


def tol_check(got, ref, rel=2e-3, abs_=1e-4, what=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs().max().item()
    lim = rel * ref.abs().max().item() + abs_
    _record(what, err, lim)
    assert err <= lim, "%s: max err %.3e > %.3e (max|ref| %.3e)" % (what, err, lim, ref.abs().max().item())
    return err


def _record(what, err, lim):
    """Append measured error / limit to gpurun_out/errors.jsonl (headroom bookkeeping, best effort)."""
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "errors.jsonl"), "a") as f:
            f.write(json.dumps({"what": what, "err": err, "lim": lim, "ratio": err / lim if lim else None}) + "\n")
    except OSError:
        pass


def pack_pw(w, bias):
    lib = _lib.load()
    n, k = w.shape
    w16 = np.ascontiguousarray(w.half().numpy()).view(np.uint16)
    wp = np.zeros(lib.lp_pw1x1_packed_elems(k, n), np.uint16)
    bp = np.zeros(lib.lp_pw1x1_packed_bias_elems(n), np.float32)
    b = None if bias is None else np.ascontiguousarray(bias.float().numpy())
    _lib.check(lib.lp_pw1x1_pack(w16.ctypes.data, None if b is None else b.ctypes.data, k, n, wp.ctypes.data,
                                 bp.ctypes.data))
    return torch.from_numpy(wp).view(torch.float16).cuda(), torch.from_numpy(bp).cuda()
