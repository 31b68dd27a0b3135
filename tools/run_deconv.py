"""Launch the fusion-deconv kernel a few times at one shape (profiling target for ncu)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from litepose_b200 import _lib

n, hw, cr, cw, co = 32, 128, 24, 16, 32
if len(sys.argv) > 1:
    n, hw, cr, cw, co = [int(v) for v in sys.argv[1:6]]
lib = _lib.load()
rs = np.random.RandomState(0)
wr = (rs.randn(cr, co, 4, 4) * 0.05).astype(np.float16).view(np.uint16)
ww = (rs.randn(cw, co, 4, 4) * 0.05).astype(np.float16).view(np.uint16)
wpk = np.zeros(lib.lp_deconv_packed_elems(cr, cw, co), np.uint16)
bpk = np.zeros(lib.lp_deconv_packed_bias_elems(co), np.float32)
_lib.check(lib.lp_deconv_pack(wr.ctypes.data, ww.ctypes.data, None, cr, cw, co, wpk.ctypes.data, bpk.ctypes.data))
wpd = torch.from_numpy(wpk).view(torch.float16).cuda()
bpd = torch.from_numpy(bpk).cuda()
a = torch.randn((n, hw, hw, cr), device="cuda").half()
b = torch.randn((n, hw, hw, cw), device="cuda").half()
out = torch.empty((n, 2 * hw, 2 * hw, co), dtype=torch.float16, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.lp_fusion_deconv_f16(a.data_ptr(), b.data_ptr(), wpd.data_ptr(), bpd.data_ptr(), out.data_ptr(),
                                        n, hw, hw, cr, cw, co, s))
torch.cuda.synchronize()
print("ok")
