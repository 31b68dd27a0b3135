"""Launch the fused output-head kernel a few times at one shape (profiling target for ncu)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from litepose_b200 import _lib

n, hw, c1, c2, co = 32, 256, 32, 16, 14
if len(sys.argv) > 1:
    n, hw, c1, c2, co = [int(v) for v in sys.argv[1:6]]
lib = _lib.load()
rs = np.random.RandomState(0)
u16 = lambda a: np.ascontiguousarray(a.astype(np.float16)).view(np.uint16)
dwc = np.zeros(lib.lp_head_fused_dw_elems(c1, c2), np.uint16)
bdc = np.zeros(dwc.size // 25, np.float32)
pwc = np.zeros(lib.lp_head_fused_pw_elems(c1, c2, co), np.uint16)
a1, a2 = u16(rs.randn(25, c1) * 0.2), u16(rs.randn(25, c2) * 0.2)
y1, y2 = np.zeros(c1, np.float32), np.zeros(c2, np.float32)
p1, p2 = u16(rs.randn(co, c1) * 0.2), u16(rs.randn(co, c2) * 0.2)
_lib.check(lib.lp_head_fused_pack(a1.ctypes.data, y1.ctypes.data, a2.ctypes.data, y2.ctypes.data, p1.ctypes.data,
                                  p2.ctypes.data, c1, c2, co, dwc.ctypes.data, bdc.ctypes.data, pwc.ctypes.data))
dwd = torch.from_numpy(dwc).view(torch.float16).cuda()
bdd = torch.from_numpy(bdc).cuda()
pwd = torch.from_numpy(pwc).view(torch.float16).cuda()
x1 = torch.randn((n, hw, hw, c1), device="cuda").half()
x2 = torch.randn((n, hw, hw, c2), device="cuda").half()
out = torch.empty((n, co, hw, hw), dtype=torch.float32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.lp_head_fused_f16(x1.data_ptr(), x2.data_ptr(), dwd.data_ptr(), bdd.data_ptr(), pwd.data_ptr(),
                                     out.data_ptr(), 1, n, hw, hw, c1, c2, co, s))
torch.cuda.synchronize()
print("ok")
