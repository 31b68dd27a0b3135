import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.nn.functional as F
from litepose_b200 import _lib
from gpu_util import from_nhwc, nhwc16, pack_pw, q16, stream
lib = _lib.load()
def run(n,h,w,ce,co,res,seed=0, wmode="rand"):
    g = torch.Generator().manual_seed(seed)
    x = q16(torch.rand(n, ce, h, w, generator=g) * 3.0)
    if wmode == "rand":
        wd = q16(torch.randn(ce, 1, 7, 7, generator=g) * 0.15)
    else:
        wd = torch.zeros(ce,1,7,7); 
        for c in range(ce): wd[c,0,(c*3)%7,(c*5)%7] = 1.0
    bd = torch.zeros(ce)
    wp = torch.zeros(co, ce)
    for c in range(min(co,ce)): wp[c, c] = 1.0      # out channel c = dw channel c
    bp = torch.zeros(co)
    mid = q16(F.relu6(F.conv2d(x, wd, bd, 1, 3, 1, ce)))
    ref = F.conv2d(mid, wp.view(co, ce, 1, 1), bp)
    wpk, bpk = pack_pw(wp, bp)
    xd = nhwc16(x)
    wdd = wd.reshape(ce, 49).t().contiguous().half().cuda()
    out = torch.full((n, h, w, co), float("nan"), dtype=torch.float16, device="cuda")
    _lib.check(lib.lp_dw7_project_f16(xd.data_ptr(), wdd.data_ptr(), bd.cuda().data_ptr(), wpk.data_ptr(), bpk.data_ptr(), None, out.data_ptr(), n, h, w, ce, co, stream()))
    torch.cuda.synchronize()
    got = from_nhwc(out)
    err = (got - ref).abs()
    print("case", (n,h,w,ce,co), wmode, "max err %.3e" % err.max().item(), "per-channel max:", [round(v,3) for v in err.amax(dim=(0,2,3)).tolist()][:16])
    if err.max() > 0.05:
        c = int(err.amax(dim=(0,2,3)).argmax())
        e = err[:, c].amax(dim=0)
        ys, xs = torch.nonzero(e > 0.05, as_tuple=True)
        print("  bad channel", c, "count", len(ys), "first", [(int(a), int(b)) for a, b in zip(ys[:8], xs[:8])], "ymin/max", int(ys.min()), int(ys.max()), "xmin/max", int(xs.min()), int(xs.max()))
for wm in ("delta", "rand"):
    run(1,16,16,32,16,False, wmode=wm)
    run(1,16,16,64,32,False, wmode=wm)
    run(1,16,16,96,48,False, wmode=wm)
    run(2,32,32,96,16,False, wmode=wm)
    run(1,16,16,96,16,False, wmode=wm)
