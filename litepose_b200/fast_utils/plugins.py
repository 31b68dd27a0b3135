"""Mirror of the reference's pybind module ``fast_utils.plugins`` (nano_demo/fast_utils/plugins.cpp:111-116) on the
sm_100a kernels lp_find_peaks_f32 / lp_assign_f32 (litepose_b200/csrc/fast_utils.cu).

Same function names, argument order and meaning, same tensor shapes/dtypes (int32 / float32, contiguous):

    find_peaks(input, tmap, threshold, window_size, max_count) -> [count, val, tag, ind]       plugins.cpp:31-64
    find_peaks_out(count, val, tag, ind, input, tmap, threshold, window_size, max_count)       plugins.cpp:9-29
    assign(cnt, val, tag, ind, joint_order, threshold, max_count) -> [num, ans]                plugins.cpp:84-108
    assign_out(num, ans, cnt, val, tag, ind, joint_order, threshold, max_count)                plugins.cpp:66-82

Differences, by construction of the GPU path:
  * tensors may live on a CUDA device (results are then returned on that device and nothing synchronises); CPU
    tensors - what the reference takes - are copied to the current CUDA device, processed there and copied back.
    There is no CPU implementation: without the CUDA library or a device the call raises;
  * dtype / contiguity / shape are checked (the reference casts data_ptr() blindly, plugins.cpp:19-24,72-78);
  * assign accepts an optional leading batch dimension (cnt [N,C], val [N,C,M] ...) and then returns num [N],
    ans [N,M,C,4]; max_count <= 32 (the reference's KM arrays hold 10 entries and overflow silently beyond that);
  * KM is capped at 4096 rounds per row: `last_status()` returns the per-image status tensor of the last assign
    on this thread (1 where the cap was hit; the reference would spin forever).
"""
import threading

import torch

from .. import _lib

_tls = threading.local()


def _dev(*tensors):
    for t in tensors:
        if t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("litepose_b200.fast_utils needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _check(t, dtype, name, ndim=None):
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if ndim is not None and t.dim() not in (ndim if isinstance(ndim, tuple) else (ndim,)):
        raise ValueError("%s: expected %s dims, got shape %s" % (name, ndim, tuple(t.shape)))


def find_peaks_out(count, val, tag, ind, input, tmap, threshold, window_size, max_count):
    _check(input, torch.float32, "input", 4)
    _check(tmap, torch.float32, "tmap", 4)
    _check(count, torch.int32, "count", 2)
    _check(val, torch.float32, "val", 3)
    _check(tag, torch.float32, "tag", 3)
    _check(ind, torch.int32, "ind", 4)
    n, c, h, w = input.shape
    m = int(max_count)
    if tuple(tmap.shape) != (n, c, h, w) or tuple(count.shape) != (n, c) or tuple(val.shape) != (n, c, m) \
            or tuple(tag.shape) != (n, c, m) or tuple(ind.shape) != (n, c, m, 2):
        raise ValueError("find_peaks_out: inconsistent shapes")
    dev = _dev(input, tmap, count, val, tag, ind)
    outs = [count, val, tag, ind]
    d_in, d_tm = input.to(dev, non_blocking=True), tmap.to(dev, non_blocking=True)
    d_out = [t.to(dev) for t in outs]
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.lp_find_peaks_f32(d_in.data_ptr(), d_tm.data_ptr(), n, c, h, w, m, float(threshold),
                                         int(window_size), d_out[0].data_ptr(), d_out[1].data_ptr(),
                                         d_out[2].data_ptr(), d_out[3].data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream), "lp_find_peaks_f32")
    for t, d in zip(outs, d_out):
        if d is not t:
            t.copy_(d)


def find_peaks(input, tmap, threshold, window_size, max_count):
    n, c = input.shape[0], input.shape[1]
    m = int(max_count)
    dev = input.device
    count = torch.zeros((n, c), dtype=torch.int32, device=dev)
    val = torch.zeros((n, c, m), dtype=torch.float32, device=dev)
    tag = torch.zeros((n, c, m), dtype=torch.float32, device=dev)
    ind = torch.zeros((n, c, m, 2), dtype=torch.int32, device=dev)
    find_peaks_out(count, val, tag, ind, input, tmap, threshold, window_size, max_count)
    return [count, val, tag, ind]


def assign_out(num, ans, cnt, val, tag, ind, joint_order, threshold, max_count):
    _check(cnt, torch.int32, "cnt", (1, 2))
    batched = cnt.dim() == 2
    _check(val, torch.float32, "val", 3 if batched else 2)
    _check(tag, torch.float32, "tag", 3 if batched else 2)
    _check(ind, torch.int32, "ind", 4 if batched else 3)
    _check(joint_order, torch.int32, "joint_order", 1)
    _check(num, torch.int32, "num", 1)
    _check(ans, torch.float32, "ans", 4 if batched else 3)
    n = cnt.shape[0] if batched else 1
    c, m = val.shape[-2], int(max_count)
    if val.shape[-1] != m or tuple(tag.shape) != tuple(val.shape) or tuple(ind.shape) != tuple(val.shape) + (2,) \
            or cnt.shape[-1] != c or joint_order.numel() < c or num.numel() != n \
            or tuple(ans.shape) != ((n, m, c, 4) if batched else (m, c, 4)):
        raise ValueError("assign_out: inconsistent shapes")
    dev = _dev(cnt, val, tag, ind, num, ans)
    outs = [num, ans]
    d = [t.to(dev, non_blocking=True) for t in (cnt, val, tag, ind, joint_order)]
    d_out = [t.to(dev) for t in outs]
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.lp_assign_f32(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(),
                                     n, c, m, float(threshold), d_out[0].data_ptr(), d_out[1].data_ptr(),
                                     status.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "lp_assign_f32")
    for t, o in zip(outs, d_out):
        if o is not t:
            t.copy_(o)
    _tls.status = status


def assign(cnt, val, tag, ind, joint_order, threshold, max_count):
    batched = cnt.dim() == 2
    c, m = val.shape[-2], int(max_count)
    dev = cnt.device
    n = cnt.shape[0] if batched else 1
    num = torch.zeros(n, dtype=torch.int32, device=dev)
    ans = torch.zeros((n, m, c, 4) if batched else (m, c, 4), dtype=torch.float32, device=dev)
    assign_out(num, ans, cnt, val, tag, ind, joint_order, threshold, max_count)
    return [num, ans]


def last_status():
    """Per-image KM status (int32, device) of the last assign/assign_out on this thread: 0 ok, 1 round cap hit."""
    return getattr(_tls, "status", None)
