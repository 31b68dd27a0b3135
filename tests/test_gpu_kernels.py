"""-m gpu: every CUDA kernel family against an fp32 CPU reference of the same op on
fp16-rounded seeded inputs (stated tolerance 2e-3*max|ref| + 1e-4), through the C ABI."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from litepose_b200 import _lib
from gpu_util import from_nhwc, nhwc16, pack_pw, q16, stream, tol_check

pytestmark = pytest.mark.gpu


def test_library_and_device():
    lib = _lib.load()
    assert lib.lp_version() >= 100
    _lib.check(lib.lp_device_check(), "device")


@pytest.mark.parametrize("h,w", [(64, 96), (32, 320), (20, 36), (16, 132)])   # vector path / >1 CTA column / generic path
@pytest.mark.parametrize("fp32_in,flip", [(True, False), (False, False), (True, True), (False, True)])
def test_stem(fp32_in, flip, h, w):
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    n = 2
    x = torch.randn(n, 3, h, w, generator=g)
    wt = q16(torch.randn(32, 3, 3, 3, generator=g) * 0.3)
    b = torch.randn(32, generator=g) * 0.1
    xin = x if fp32_in else q16(x)
    src = torch.flip(xin, [3]) if flip else xin
    ref = F.relu6(F.conv2d(src, wt, b, 2, 1))
    xd = (xin if fp32_in else xin.half()).cuda().contiguous()
    y = torch.empty(n, h // 2, w // 2, 32, dtype=torch.float16, device="cuda")
    wd = wt.reshape(32, 27).half().cuda()
    _lib.check(lib.lp_stem_conv3x3_s2(xd.data_ptr(), 1 if fp32_in else 0, 1 if flip else 0, wd.data_ptr(),
                                      b.cuda().data_ptr(), y.data_ptr(), n, h, w, stream()), "stem")
    torch.cuda.synchronize()
    tol_check(from_nhwc(y), ref, what="stem")


@pytest.mark.parametrize("k,s,c,h,w,act", [
    (7, 1, 96, 32, 32, 2), (7, 2, 96, 64, 64, 2), (7, 1, 48, 40, 24, 2), (7, 1, 720, 16, 16, 2),
    (7, 2, 144, 34 * 2, 18 * 2, 2), (5, 1, 24, 64, 64, 1), (5, 1, 40, 36, 68, 1), (3, 1, 32, 64, 64, 2),
    (3, 2, 16, 32, 32, 0), (7, 1, 8, 8, 8, 2),
])
@pytest.mark.parametrize("prec", [0, 1, 2])
def test_dwconv(k, s, c, h, w, act, prec):
    lib = _lib.load()
    lib.lp_set_dw_precision(prec)
    g = torch.Generator().manual_seed(k * 100 + c)
    n = 2
    x = q16(torch.randn(n, c, h, w, generator=g))
    wt = q16(torch.randn(c, 1, k, k, generator=g) * 0.2)
    b = torch.randn(c, generator=g) * 0.1
    ref = F.conv2d(x, wt, b, s, k // 2, 1, c)
    ref = F.relu6(ref) if act == 2 else (F.relu(ref) if act == 1 else ref)
    xd = nhwc16(x)
    wd = wt.reshape(c, k * k).t().contiguous().half().cuda()
    y = torch.full((n, h // s, w // s, c), float("nan"), dtype=torch.float16, device="cuda")
    _lib.check(lib.lp_dwconv_f16(xd.data_ptr(), wd.data_ptr(), b.cuda().data_ptr(), y.data_ptr(), n, c, h, w, k, s,
                                 act, stream()), "dwconv")
    torch.cuda.synchronize()
    lib.lp_set_dw_precision(-1)           # back to the per-kernel-size default
    tol_check(from_nhwc(y), ref, what="dwconv k%d s%d c%d prec%d" % (k, s, c, prec))


@pytest.mark.parametrize("m,k,n,act,res", [
    (128, 16, 96, 2, False), (1000, 96, 16, 0, True), (4096, 32, 192, 2, False), (640, 192, 32, 0, False),
    (2048, 120, 720, 2, False), (2048, 720, 120, 0, True), (300, 24, 144, 2, False), (260, 432, 72, 0, True),
    (512, 160, 960, 2, False), (64, 960, 160, 0, False), (4096, 32, 16, 0, False), (20000, 48, 288, 2, False),
])
def test_pw1x1(m, k, n, act, res):
    lib = _lib.load()
    g = torch.Generator().manual_seed(m + k + n)
    a = q16(torch.randn(m, k, generator=g))
    wt = q16(torch.randn(n, k, generator=g) / (k ** 0.5))
    b = torch.randn(n, generator=g) * 0.1
    r = q16(torch.randn(m, n, generator=g)) if res else None
    ref = a @ wt.t() + b
    ref = F.relu6(ref) if act == 2 else (F.relu(ref) if act == 1 else ref)
    if res:
        ref = ref + r
    wp, bp = pack_pw(wt, b)
    ad = a.half().cuda()
    rd = r.half().cuda() if res else None
    out = torch.full((m, n), float("nan"), dtype=torch.float16, device="cuda")
    _lib.check(lib.lp_pw1x1_f16(ad.data_ptr(), wp.data_ptr(), bp.data_ptr(), rd.data_ptr() if res else None,
                                out.data_ptr(), m, k, n, act, stream()), "pw1x1")
    torch.cuda.synchronize()
    tol_check(out, ref, what="pw %dx%dx%d" % (m, k, n))


@pytest.mark.parametrize("n,h,w,cr,cw,co", [
    (2, 16, 16, 80, 48, 16), (1, 32, 32, 120, 48, 32), (2, 20, 12, 160, 96, 64), (1, 64, 64, 24, 16, 32),
    (1, 28, 28, 32, 32, 24), (3, 8, 8, 16, 24, 24),
    # wide maps: row-streaming kernel (strips of 128 px; partial strips; odd chunk counts; long per-CTA row chains)
    (2, 8, 128, 24, 16, 32), (1, 5, 160, 24, 16, 32), (2, 12, 96, 32, 32, 24), (1, 7, 128, 16, 8, 16),
    (2, 3, 256, 8, 8, 8), (8, 128, 128, 24, 16, 32), (1, 1, 128, 40, 24, 40),
])
def test_fusion_deconv(n, h, w, cr, cw, co):
    lib = _lib.load()
    g = torch.Generator().manual_seed(cr + cw + co)
    xr = q16(torch.randn(n, cr, h, w, generator=g))
    xw = q16(torch.randn(n, cw, h, w, generator=g))
    wr = q16(torch.randn(cr, co, 4, 4, generator=g) / (cr ** 0.5))
    ww = q16(torch.randn(cw, co, 4, 4, generator=g) / (cw ** 0.5))
    b = torch.randn(co, generator=g) * 0.1
    ref = F.relu(F.conv_transpose2d(xr, wr, None, 2, 1) + F.conv_transpose2d(xw, ww, None, 2, 1) + b.view(1, -1, 1, 1))
    wp = np.zeros(lib.lp_deconv_packed_elems(cr, cw, co), np.uint16)
    bp = np.zeros(lib.lp_deconv_packed_bias_elems(co), np.float32)
    a16 = np.ascontiguousarray(wr.half().numpy()).view(np.uint16)
    c16 = np.ascontiguousarray(ww.half().numpy()).view(np.uint16)
    bb = np.ascontiguousarray(b.numpy())
    _lib.check(lib.lp_deconv_pack(a16.ctypes.data, c16.ctypes.data, bb.ctypes.data, cr, cw, co, wp.ctypes.data,
                                  bp.ctypes.data))
    wd = torch.from_numpy(wp).view(torch.float16).cuda()
    bd = torch.from_numpy(bp).cuda()
    out = torch.full((n, 2 * h, 2 * w, co), float("nan"), dtype=torch.float16, device="cuda")
    r_, w_ = nhwc16(xr), nhwc16(xw)
    _lib.check(lib.lp_fusion_deconv_f16(r_.data_ptr(), w_.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(),
                                        n, h, w, cr, cw, co, stream()), "deconv")
    torch.cuda.synchronize()
    tol_check(from_nhwc(out), ref, what="deconv")


@pytest.mark.parametrize("n,h,w,c1,c2,co,fp32", [
    (2, 32, 32, 24, 16, 28, True), (1, 64, 64, 32, 16, 14, True), (1, 40, 24, 40, 24, 28, False),
    (2, 16, 16, 24, 16, 34, True),
])
def test_head(n, h, w, c1, c2, co, fp32):
    lib = _lib.load()
    g = torch.Generator().manual_seed(c1 + c2 + co)
    a1 = q16(torch.randn(n, c1, h, w, generator=g))
    a2 = q16(torch.randn(n, c2, h, w, generator=g))
    w1 = q16(torch.randn(co, c1, generator=g) / (c1 ** 0.5))
    w2 = q16(torch.randn(co, c2, generator=g) / (c2 ** 0.5))
    ref = F.conv2d(a1, w1.view(co, c1, 1, 1)) + F.conv2d(a2, w2.view(co, c2, 1, 1))
    wp = np.zeros(lib.lp_head_packed_elems(c1, c2, co), np.uint16)
    x16 = np.ascontiguousarray(w1.half().numpy()).view(np.uint16)
    y16 = np.ascontiguousarray(w2.half().numpy()).view(np.uint16)
    _lib.check(lib.lp_head_pack(x16.ctypes.data, y16.ctypes.data, c1, c2, co, wp.ctypes.data))
    wd = torch.from_numpy(wp).view(torch.float16).cuda()
    out = torch.full((n, co, h, w), float("nan"), dtype=torch.float32 if fp32 else torch.float16, device="cuda")
    d1, d2 = nhwc16(a1), nhwc16(a2)
    _lib.check(lib.lp_head_pw_dual_f16(d1.data_ptr(), d2.data_ptr(), wd.data_ptr(), out.data_ptr(), 1 if fp32 else 0,
                                       n, h, w, c1, c2, co, stream()), "head")
    torch.cuda.synchronize()
    tol_check(out, ref, what="head")


def test_error_codes():
    lib = _lib.load()
    rc = lib.lp_dwconv_f16(None, None, None, None, 1, 8, 8, 8, 7, 1, 0, None)
    assert rc == 1 and b"null" in lib.lp_last_error()
    x = torch.zeros(64, dtype=torch.float16, device="cuda")
    rc = lib.lp_dwconv_f16(x.data_ptr(), x.data_ptr(), None, x.data_ptr(), 1, 12, 8, 8, 7, 1, 0, None)
    assert rc == 1
    rc = lib.lp_pw1x1_f16(x.data_ptr(), x.data_ptr(), None, None, x.data_ptr(), 8, 12, 8, 0, None)
    assert rc == 1


@pytest.mark.parametrize("n,h,w,ce,co,res", [
    (2, 32, 32, 96, 16, True), (1, 16, 16, 96, 16, False), (2, 48, 32, 192, 32, True), (1, 32, 32, 288, 48, True),
    (1, 32, 32, 720, 120, True), (2, 16, 16, 144, 24, True), (1, 20, 40, 432, 72, False), (1, 32, 32, 288, 120, False),
    (3, 64, 64, 96, 16, True),
    # more tiles than SMs: several tiles per persistent CTA (operand / accumulator double buffering, alternating slab
    # groups for odd slab counts, deferred epilogue, single-buffer weight ring for wide projections)
    (8, 128, 128, 96, 16, True), (5, 96, 112, 160, 32, True), (36, 32, 32, 288, 48, True), (24, 48, 48, 720, 120, True),
    (20, 48, 48, 48, 8, False), (6, 80, 80, 32, 16, True), (1, 16, 16, 960, 160, True),
])
def test_dw7_project_fused(n, h, w, ce, co, res):
    """fused depthwise-7x7 + projection (+residual) against the unfused fp32 reference"""
    lib = _lib.load()
    g = torch.Generator().manual_seed(ce + co + h)
    x = q16(torch.rand(n, ce, h, w, generator=g) * 3.0)
    wd = q16(torch.randn(ce, 1, 7, 7, generator=g) * 0.15)
    bd = torch.randn(ce, generator=g) * 0.1
    wp = q16(torch.randn(co, ce, generator=g) / (ce ** 0.5))
    bp = torch.randn(co, generator=g) * 0.1
    r = q16(torch.randn(n, co, h, w, generator=g)) if res else None
    mid = q16(F.relu6(F.conv2d(x, wd, bd, 1, 3, 1, ce)))
    ref = F.conv2d(mid, wp.view(co, ce, 1, 1), bp)
    if res:
        ref = ref + r
    wpk, bpk = pack_pw(wp, bp)
    xd = nhwc16(x)
    wdd = wd.reshape(ce, 49).t().contiguous().half().cuda()
    rd = nhwc16(r) if res else None
    out = torch.full((n, h, w, co), float("nan"), dtype=torch.float16, device="cuda")
    _lib.check(lib.lp_dw7_project_f16(xd.data_ptr(), wdd.data_ptr(), bd.cuda().data_ptr(), wpk.data_ptr(),
                                      bpk.data_ptr(), rd.data_ptr() if res else None, out.data_ptr(), n, h, w, ce, co,
                                      stream()), "dw7_project")
    torch.cuda.synchronize()
    tol_check(from_nhwc(out), ref, what="dw7_project ce%d co%d" % (ce, co))


@pytest.mark.parametrize("n,h,w,c1,c2,co,fp32", [
    (2, 32, 32, 24, 16, 28, True), (1, 64, 64, 32, 16, 14, True), (1, 48, 16, 40, 24, 28, False),
    (2, 16, 16, 64, 24, 34, True), (1, 20, 36, 24, 16, 28, True),
    (4, 144, 160, 32, 16, 14, True), (6, 96, 96, 24, 16, 28, True), (5, 80, 96, 40, 24, 28, False),     # many tiles per CTA
])
def test_head_fused(n, h, w, c1, c2, co, fp32):
    """both SepConv2d heads of a level (dw5+BN+ReLU -> 1x1, two branches summed) in one kernel"""
    lib = _lib.load()
    g = torch.Generator().manual_seed(c1 * 7 + c2 + co)
    x1 = q16(torch.randn(n, c1, h, w, generator=g))
    x2 = q16(torch.randn(n, c2, h, w, generator=g))
    d1 = q16(torch.randn(c1, 1, 5, 5, generator=g) * 0.2)
    d2 = q16(torch.randn(c2, 1, 5, 5, generator=g) * 0.2)
    b1, b2 = torch.randn(c1, generator=g) * 0.1, torch.randn(c2, generator=g) * 0.1
    w1 = q16(torch.randn(co, c1, generator=g) / (c1 ** 0.5))
    w2 = q16(torch.randn(co, c2, generator=g) / (c2 ** 0.5))
    m1 = q16(F.relu(F.conv2d(x1, d1, b1, 1, 2, 1, c1)))
    m2 = q16(F.relu(F.conv2d(x2, d2, b2, 1, 2, 1, c2)))
    ref = F.conv2d(m1, w1.view(co, c1, 1, 1)) + F.conv2d(m2, w2.view(co, c2, 1, 1))

    def u16(t):
        return np.ascontiguousarray(t.half().numpy()).view(np.uint16)

    dwc = np.zeros(lib.lp_head_fused_dw_elems(c1, c2), np.uint16)
    bdc = np.zeros(dwc.size // 25, np.float32)
    pwc = np.zeros(lib.lp_head_fused_pw_elems(c1, c2, co), np.uint16)
    a1, a2 = u16(d1.reshape(c1, 25).t().contiguous()), u16(d2.reshape(c2, 25).t().contiguous())
    y1, y2 = np.ascontiguousarray(b1.numpy()), np.ascontiguousarray(b2.numpy())
    p1, p2 = u16(w1), u16(w2)
    _lib.check(lib.lp_head_fused_pack(a1.ctypes.data, y1.ctypes.data, a2.ctypes.data, y2.ctypes.data, p1.ctypes.data,
                                      p2.ctypes.data, c1, c2, co, dwc.ctypes.data, bdc.ctypes.data, pwc.ctypes.data))
    dwd = torch.from_numpy(dwc).view(torch.float16).cuda()
    bdd = torch.from_numpy(bdc).cuda()
    pwd = torch.from_numpy(pwc).view(torch.float16).cuda()
    out = torch.full((n, co, h, w), float("nan"), dtype=torch.float32 if fp32 else torch.float16, device="cuda")
    s1, s2 = nhwc16(x1), nhwc16(x2)
    _lib.check(lib.lp_head_fused_f16(s1.data_ptr(), s2.data_ptr(), dwd.data_ptr(), bdd.data_ptr(), pwd.data_ptr(),
                                     out.data_ptr(), 1 if fp32 else 0, n, h, w, c1, c2, co, stream()), "head_fused")
    torch.cuda.synchronize()
    tol_check(out, ref, what="head_fused c1 %d c2 %d co %d" % (c1, c2, co))


@pytest.mark.parametrize("n,h,w,cin,ce,co,res", [
    (1, 16, 16, 16, 96, 16, True), (2, 32, 32, 16, 96, 16, True), (1, 32, 48, 32, 192, 32, True),
    (2, 20, 36, 24, 144, 24, True),            # ragged tiles, K padded 24 -> 32
    (1, 16, 16, 16, 32, 16, False),            # a single slab
    (1, 32, 32, 32, 64, 48, False),            # two slabs, no identity
    (1, 16, 16, 64, 160, 64, True),            # widest supported input / output
    # more tiles than SMs: several tiles per persistent CTA (X buffer hand-over, slot phases across tiles)
    (8, 128, 128, 16, 96, 16, True), (6, 80, 96, 32, 192, 32, True), (3, 112, 112, 24, 144, 24, True),
    # streaming mode (weights through two-slot rings): stage-2-class blocks, odd and even slab counts, several tiles per CTA
    (1, 32, 32, 48, 288, 48, True), (2, 16, 16, 48, 288, 48, False), (40, 32, 32, 48, 288, 48, True), (3, 96, 96, 48, 288, 48, True),
    (2, 32, 32, 64, 384, 64, True), (1, 48, 48, 40, 320, 56, False),
])
def test_block_s1_fused(n, h, w, cin, ce, co, res):
    """whole stride-1 InvBottleneck in one kernel (expand on tensor cores -> slab -> depthwise -> projection -> identity)
    against the unfused fp32 reference with fp16 storage of the two intermediates (what the unfused kernels compute)"""
    lib = _lib.load()
    assert lib.lp_block_s1_supported(cin, ce, co) == 1
    g = torch.Generator().manual_seed(cin * 3 + ce + co + h)
    x = q16(torch.randn(n, cin, h, w, generator=g))
    we = q16(torch.randn(ce, cin, generator=g) / (cin ** 0.5))
    be = torch.randn(ce, generator=g) * 0.2
    wd = q16(torch.randn(ce, 1, 7, 7, generator=g) * 0.1)
    bd = torch.randn(ce, generator=g) * 0.1
    wp = q16(torch.randn(co, ce, generator=g) / (ce ** 0.5))
    bp = torch.randn(co, generator=g) * 0.1
    e = q16(F.relu6(F.conv2d(x, we.view(ce, cin, 1, 1), be)))
    mid = q16(F.relu6(F.conv2d(e, wd, bd, 1, 3, 1, ce)))
    ref = F.conv2d(mid, wp.view(co, ce, 1, 1), bp)
    if res:
        ref = ref + x
    wpk, bpk = pack_pw(wp, bp)
    we16 = np.ascontiguousarray(we.half().numpy()).view(np.uint16)
    wek = np.zeros(lib.lp_block_s1_wexp_elems(cin, ce), np.uint16)
    _lib.check(lib.lp_block_s1_pack_wexp(we16.ctypes.data, cin, ce, wek.ctypes.data))
    wed = torch.from_numpy(wek).view(torch.float16).cuda()
    xd = nhwc16(x)
    wdd = wd.reshape(ce, 49).t().contiguous().half().cuda()
    out = torch.full((n, h, w, co), float("nan"), dtype=torch.float16, device="cuda")
    bed, bdd = be.cuda(), bd.cuda()      # keep the device copies alive (two temporaries would alias in the allocator)
    for _ in range(2):       # second launch: no state leaks between launches
        _lib.check(lib.lp_block_s1_f16(xd.data_ptr(), wed.data_ptr(), bed.data_ptr(), wdd.data_ptr(),
                                       bdd.data_ptr(), wpk.data_ptr(), bpk.data_ptr(), 1 if res else 0,
                                       out.data_ptr(), n, h, w, cin, ce, co, stream()), "block_s1")
    torch.cuda.synchronize()
    tol_check(from_nhwc(out), ref, what="block_s1 cin%d ce%d co%d" % (cin, ce, co))


def test_block_s1_unsupported_shapes():
    lib = _lib.load()
    assert lib.lp_block_s1_supported(72, 432, 72) == 0      # Cin > 64
    assert lib.lp_block_s1_supported(120, 720, 120) == 0    # Cin > 64
    x = torch.zeros(1, 16, 16, 120, dtype=torch.float16, device="cuda")
    rc = lib.lp_block_s1_f16(x.data_ptr(), x.data_ptr(), None, x.data_ptr(), None, x.data_ptr(), None, 0, x.data_ptr(),
                             1, 16, 16, 120, 720, 120, stream())
    assert rc != 0 and b"unsupported shape" in lib.lp_last_error()


@pytest.mark.parametrize("n,h,w,c0,fp32_in,flip", [
    (2, 64, 64, 16, True, False), (1, 128, 96, 16, False, True), (2, 48, 80, 24, False, False),
    (1, 32, 36, 16, True, True),                 # ragged tiles (Wo = 18, Ho = 16)
    (3, 256, 256, 16, False, False),             # many tiles per persistent CTA
    (1, 512, 512, 16, False, True),
])
def test_stem_fused(n, h, w, c0, fp32_in, flip):
    """conv3x3 s2 + BN + ReLU6 -> dw3x3 + BN + ReLU6 -> 1x1 + BN in one kernel vs the fp32 reference with fp16 storage of
    the two intermediates (what the three unfused kernels compute); reference lib/models/pose_mobilenet.py:36-41"""
    lib = _lib.load()
    assert lib.lp_stem_fused_supported(h, w, c0) == 1
    g = torch.Generator().manual_seed(h + w + c0)
    xin = q16(torch.randn(n, 3, h, w, generator=g))
    w1 = q16(torch.randn(32, 3, 3, 3, generator=g) * 0.3)
    b1 = torch.randn(32, generator=g) * 0.2
    wd = q16(torch.randn(32, 1, 3, 3, generator=g) * 0.3)
    bd = torch.randn(32, generator=g) * 0.2
    w3 = q16(torch.randn(c0, 32, generator=g) / 32 ** 0.5)
    b3 = torch.randn(c0, generator=g) * 0.1
    src = torch.flip(xin, [3]) if flip else xin
    a0 = q16(F.relu6(F.conv2d(src, w1, b1, 2, 1)))
    a1 = q16(F.relu6(F.conv2d(a0, wd, bd, 1, 1, 1, 32)))
    ref = F.conv2d(a1, w3.view(c0, 32, 1, 1), b3)
    w1p = torch.zeros(32, 64, dtype=torch.float16)
    w1p[:, :27] = w1.reshape(32, 27).half()
    w3k, b3k = pack_pw(w3, b3)
    xd = (xin if fp32_in else xin.half()).cuda().contiguous()
    w1d, b1d = w1p.cuda(), b1.cuda()
    wdd, bdd = wd.reshape(32, 9).t().contiguous().half().cuda(), bd.cuda()
    y = torch.full((n, h // 2, w // 2, c0), float("nan"), dtype=torch.float16, device="cuda")
    for _ in range(2):
        _lib.check(lib.lp_stem_fused_f16(xd.data_ptr(), 1 if fp32_in else 0, 1 if flip else 0, w1d.data_ptr(), b1d.data_ptr(),
                                         wdd.data_ptr(), bdd.data_ptr(), w3k.data_ptr(), b3k.data_ptr(), y.data_ptr(), n, h, w,
                                         c0, stream()), "stem_fused")
    torch.cuda.synchronize()
    tol_check(from_nhwc(y), ref, what="stem_fused c0=%d %dx%d" % (c0, h, w))
