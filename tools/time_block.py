"""Times the block-level kernels at the stage shapes of LitePose-S @512 batch 32 (CUDA events, L2 flushed):
fused block kernel vs expansion GEMM + fused depthwise/projection kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from litepose_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=12, skip=3):
    ts = []
    for i in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= skip: ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts)

def pack_pw(w, K, N):
    w16 = np.ascontiguousarray(w.astype(np.float16)).view(np.uint16)
    wp = np.zeros(lib.lp_pw1x1_packed_elems(K, N), np.uint16); bp = np.zeros(lib.lp_pw1x1_packed_bias_elems(N), np.float32)
    _lib.check(lib.lp_pw1x1_pack(w16.ctypes.data, None, K, N, wp.ctypes.data, bp.ctypes.data))
    return torch.from_numpy(wp).view(torch.float16).to(dev), torch.from_numpy(bp).to(dev)

out = {}
s = torch.cuda.current_stream().cuda_stream
for (n, hw, cin, ce, co) in ((32, 128, 16, 96, 16), (32, 64, 32, 192, 32)):
    rs = np.random.RandomState(0)
    x = torch.randn((n, hw, hw, cin), device=dev).half()
    we = (rs.randn(ce, cin) / cin ** 0.5).astype(np.float32)
    wek = np.zeros(lib.lp_block_s1_wexp_elems(cin, ce), np.uint16)
    w16 = np.ascontiguousarray(we.astype(np.float16)).view(np.uint16)
    _lib.check(lib.lp_block_s1_pack_wexp(w16.ctypes.data, cin, ce, wek.ctypes.data))
    wed = torch.from_numpy(wek).view(torch.float16).to(dev)
    be = torch.zeros(ce, device=dev); bd = torch.zeros(ce, device=dev)
    wd = (torch.randn((49, ce), device=dev) * 0.1).half()
    wpd, bpd = pack_pw(rs.randn(co, ce) / ce ** 0.5, ce, co)
    wexp, bexp = pack_pw(we, cin, ce)
    e = torch.empty((n, hw, hw, ce), dtype=torch.float16, device=dev)
    o = torch.empty((n, hw, hw, co), dtype=torch.float16, device=dev)
    t_blk = timeit(lambda: _lib.check(lib.lp_block_s1_f16(x.data_ptr(), wed.data_ptr(), be.data_ptr(), wd.data_ptr(), bd.data_ptr(),
                                                           wpd.data_ptr(), bpd.data_ptr(), 1, o.data_ptr(), n, hw, hw, cin, ce, co, s)))
    t_exp = timeit(lambda: _lib.check(lib.lp_pw1x1_f16(x.data_ptr(), wexp.data_ptr(), bexp.data_ptr(), None, e.data_ptr(), n * hw * hw, cin, ce, 2, s)))
    t_dwp = timeit(lambda: _lib.check(lib.lp_dw7_project_f16(e.data_ptr(), wd.data_ptr(), bd.data_ptr(), wpd.data_ptr(), bpd.data_ptr(),
                                                              x.data_ptr(), o.data_ptr(), n, hw, hw, ce, co, s)))
    out["%dx%dx%dx%d->%d->%d" % (n, hw, hw, cin, ce, co)] = {"block_us": t_blk, "expand_us": t_exp, "dw_project_us": t_dwp}
print(json.dumps({"skew_ns": os.environ.get("LP_BLOCK_SKEW_NS", "0"), "shapes": out}))
