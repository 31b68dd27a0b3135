"""Error-budget emulation behind DESIGN.md section 5 (test infrastructure, CPU only; not collected by pytest):

    python tests/emulate_dw_precision.py [ARCH SIZE ...]        # default: XS 128, S 128, S 256

Runs the BN-folded network (oracle.model_ref.fold_bn) with fp16 storage of every activation and the stride-1 7x7 depthwise
accumulated (a) in fp32 [mode None = what round 1 computed], (b) as fp16 row chains summed in fp32 ['rows32' = LP_DW_PREC=1],
(c) as fp16 row chains folded by an fp16 tree ['rows' ~ the shipped grouped chains], (d) as ONE 49-tap fp16 chain ['full'],
and prints max|out - fp32 oracle| / (2e-3 * max|ref| + 1e-4) for the two network outputs.  Measured here (round 2):
XS/S 0.25-0.30 -> 0.28-0.35 even for (d); M 0.58 -> 0.59, L 0.44 -> 0.36: the budget is dominated by the fp16 activation
storage, which the reference's own fp16 path shares."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from litepose_b200 import synth
from litepose_b200.config import get_arch, get_cfg
from litepose_b200.lib.models.pose_mobilenet import get_pose_net
from oracle import model_ref

def h(t): return t.half().float()

def dw_emul(x, w, b, mode):
    """x [N,C,H,W] fp32 holding fp16 values, w [C,1,7,7] (fp16 values), b [C] fp32. mode: 'rows' = fp16 fma chain per row, then fp16 pairwise tree over rows + bias;
    'full' = one fp16 chain of 49; 'rows32' = fp16 row chains, fp32 sum of rows (PREC=1)"""
    N, C, H, W = x.shape
    xp = F.pad(x, (3, 3, 3, 3)).half()
    wh = w.half()
    rows = []
    acc_full = None
    for ky in range(7):
        acc = None
        for kx in range(7):
            xs = xp[:, :, ky:ky + H, kx:kx + W]
            ws = wh[:, 0, ky, kx].view(1, C, 1, 1)
            prod32 = xs.float() * ws.float()
            if mode == 'full':
                acc_full = prod32.half() if acc_full is None else (acc_full.float() + prod32).half()   # fma: single rounding
            else:
                acc = prod32.half() if acc is None else (acc.float() + prod32).half()
        rows.append(acc)
    if mode == 'full':
        return (acc_full.float() + b.view(1, C, 1, 1)).half().float()
    if mode == 'rows32':
        s = sum(r.float() for r in rows) + b.view(1, C, 1, 1)
        return s
    # fp16 tree: ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+bias))
    bh = b.view(1, C, 1, 1).half().expand_as(rows[0])
    a = (rows[0].float() + rows[1].float()).half(); b2 = (rows[2].float() + rows[3].float()).half()
    c = (rows[4].float() + rows[5].float()).half(); d = (rows[6].float() + bh.float()).half()
    e = (a.float() + b2.float()).half(); f = (c.float() + d.float()).half()
    return (e.float() + f.float()).half().float()

def forward(fp, arch, x, mode):
    q = h
    def conv(name, x, stride=1, groups=1, act=None):
        w, b = fp[name]
        if mode and groups > 1 and w.shape[-1] == 7 and stride == 1:
            y = dw_emul(x, q(w), b, mode)
        else:
            y = F.conv2d(x, q(w), b, stride, w.shape[-1] // 2, 1, groups)
        if act == "relu6": y = F.relu6(y)
        elif act == "relu": y = F.relu(y)
        return y
    x = q(x.float())
    x = q(conv("first.0", x, 2, 1, "relu6")); x = q(conv("first.1", x, 1, x.shape[1], "relu6")); x = q(conv("first.2", x))
    x_list = [x]
    for si, st in enumerate(arch["backbone_setting"]):
        for bi in range(st["num_blocks"]):
            p = "stage.%d.%d." % (si, bi); stride = st["stride"] if bi == 0 else 1
            inp = x
            y = q(conv(p + "inv", x, 1, 1, "relu6"))
            y = q(conv(p + "depth_conv", y, stride, y.shape[1], "relu6"))
            y = conv(p + "point_conv", y)
            if stride == 1 and inp.shape[1] == y.shape[1]: y = y + inp
            x = q(y)
        x_list.append(x)
    outs = []
    refined, raw = x_list[-1], x_list[-2]
    for i in range(3):
        wr, b = fp["deconv_refined.%d" % i]; ww, _ = fp["deconv_raw.%d" % i]
        y = F.conv_transpose2d(refined, q(wr), None, 2, 1) + F.conv_transpose2d(raw, q(ww), None, 2, 1)
        refined = q(F.relu(y + b.view(1, -1, 1, 1))); raw = x_list[-i - 3]
        if i > 0:
            o = 0
            for nm, src in (("final_refined", refined), ("final_raw", raw)):
                p = "%s.%d.conv." % (nm, i - 1)
                t = q(conv(p + "dw", src, 1, src.shape[1], "relu"))
                o = o + F.conv2d(t, q(fp[p + "pw"][0]))
            outs.append(o)
    return outs

torch.set_num_threads(8)
args = sys.argv[1:]
cases = [(args[i], int(args[i + 1])) for i in range(0, len(args) - 1, 2)] or [("XS", 128), ("S", 128), ("S", 256)]
for name, size in cases:
    cfg = get_cfg(input_size=size); arch = get_arch(name)
    torch.manual_seed(0)
    model = synth.randomize_bn_(get_pose_net(cfg, False, arch), 1).eval()
    sd = model.state_dict()
    x = synth.make_frames(1, size, seed=11)
    with torch.no_grad():
        ref = model_ref.forward(sd, arch, x)
        fp = model_ref.fold_bn(sd, arch)
        for mode in (None, 'rows32', 'rows', 'full'):
            o = forward(fp, arch, x, mode)
            r = [float((a - b).abs().max() / (2e-3 * b.abs().max() + 1e-4)) for a, b in zip(o, ref)]
            print(name, size, mode, ["%.3f" % v for v in r], flush=True)
