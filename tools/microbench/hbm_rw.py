"""HBM bandwidth by access mix on this GPU (CUDA events, best of 10): pure write (fill), pure read (sum), copy.
The expansion GEMMs of the network are almost pure writes (16 -> 96 channels), so their roofline is the WRITE figure."""
import json
import torch

dev = torch.device("cuda", 0)
n = 1 << 29   # 512 Mi fp16 = 1 GiB
a = torch.empty(n, dtype=torch.float16, device=dev)
b = torch.empty(n, dtype=torch.float16, device=dev)
a.fill_(1.0)
out = {}


def best(fn, byts, reps=10):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return byts / min(ts) / 1e9


out["write_fill_gbs"] = best(lambda: b.fill_(2.0), n * 2)
out["write_zero_gbs"] = best(lambda: b.zero_(), n * 2)
out["read_sum_gbs"] = best(lambda: a.sum(), n * 2)
out["copy_gbs_read_plus_write"] = best(lambda: b.copy_(a), n * 4)
# 1:6 read:write mix like an expansion GEMM: out[6n] = f(in[n])
c = a[: n // 6]
d = b[: (n // 6) * 6].view(6, -1)
out["expand_1to6_gbs"] = best(lambda: torch.mul(c.unsqueeze(0), 2.0, out=d), (n // 6) * 7 * 2)
print(json.dumps(out))
