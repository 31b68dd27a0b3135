import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.nn.functional as F
from litepose_b200 import _lib
from gpu_util import from_nhwc, nhwc16, pack_pw, q16, stream
lib = _lib.load()
torch.set_printoptions(linewidth=250, precision=3, sci_mode=False)

def run(n, h, w, cin, ce, co, mode, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = q16(torch.rand(n, cin, h, w, generator=g) * 2.0)
    be = torch.zeros(ce); bd = torch.zeros(ce); bp = torch.zeros(co)
    if mode == "delta":
        we = torch.zeros(ce, cin)
        for c in range(ce): we[c, c % cin] = 1.0
        wd = torch.zeros(ce, 1, 7, 7); wd[:, 0, 3, 3] = 1.0
        wp = torch.zeros(co, ce)
        for c in range(min(co, ce)): wp[c, c] = 1.0
    elif mode == "shift":
        we = torch.zeros(ce, cin)
        for c in range(ce): we[c, c % cin] = 1.0
        wd = torch.zeros(ce, 1, 7, 7)
        for c in range(ce): wd[c, 0, (c * 3) % 7, (c * 5) % 7] = 1.0
        wp = torch.zeros(co, ce)
        for c in range(min(co, ce)): wp[c, c] = 1.0
    else:
        we = q16(torch.randn(ce, cin, generator=g) / cin ** 0.5)
        wd = q16(torch.randn(ce, 1, 7, 7, generator=g) * 0.1)
        wp = q16(torch.randn(co, ce, generator=g) / ce ** 0.5)
        be = torch.randn(ce, generator=g) * 0.2
    e = q16(F.relu6(F.conv2d(x, we.view(ce, cin, 1, 1), be)))
    mid = q16(F.relu6(F.conv2d(e, wd, bd, 1, 3, 1, ce)))
    ref = F.conv2d(mid, wp.view(co, ce, 1, 1), bp)
    wpk, bpk = pack_pw(wp, bp)
    we16 = np.ascontiguousarray(we.half().numpy()).view(np.uint16)
    wek = np.zeros(lib.lp_block_s1_wexp_elems(cin, ce), np.uint16)
    _lib.check(lib.lp_block_s1_pack_wexp(we16.ctypes.data, cin, ce, wek.ctypes.data))
    wed = torch.from_numpy(wek).view(torch.float16).cuda()
    xd = nhwc16(x)
    wdd = wd.reshape(ce, 49).t().contiguous().half().cuda()
    out = torch.full((n, h, w, co), float("nan"), dtype=torch.float16, device="cuda")
    bed, bdd = be.cuda(), bd.cuda()
    _lib.check(lib.lp_block_s1_f16(xd.data_ptr(), wed.data_ptr(), bed.data_ptr(), wdd.data_ptr(), bdd.data_ptr(),
                                   wpk.data_ptr(), bpk.data_ptr(), 0, out.data_ptr(), n, h, w, cin, ce, co, stream()))
    torch.cuda.synchronize()
    got = from_nhwc(out)
    err = (got - ref).abs()
    print("case", (n, h, w, cin, ce, co), mode, "max err %.3e max ref %.3f nan %d" % (err.nan_to_num(9).max().item(), ref.abs().max().item(), int(torch.isnan(got).sum())))
    pc = err.nan_to_num(9).amax(dim=(0, 2, 3))
    print("  per-channel max err:", [round(v, 3) for v in pc.tolist()])
    if err.nan_to_num(9).max() > 0.02:
        c = int(pc.argmax())
        em = err[0, c].nan_to_num(9)
        print("  worst channel", c, "error map rows (x = >0.02):")
        for y in range(min(h, 32)):
            print("   ", "".join("x" if em[y, xx] > 0.02 else "." for xx in range(min(w, 48))))
        ys, xs = torch.nonzero(em > 0.02, as_tuple=True)
        if len(ys):
            y0, x0 = int(ys[0]), int(xs[0])
            print("  first bad (y,x)=", (y0, x0), "got", got[0, c, y0, x0].item(), "ref", ref[0, c, y0, x0].item())

for mode in ("delta", "shift", "rand"):
    run(1, 16, 16, 16, 32, 16, mode)
    run(1, 16, 16, 16, 96, 16, mode)
    run(1, 32, 32, 32, 64, 32, mode)
