"""Times the EdgeConv kernels at the C2 shapes (B=64, N=4096, k=20) in fp32 and tf32 modes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops

def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

B, N, k, C = 64, 4096, 20, 128
M = B * N
pq = torch.randn(M, 2 * C, device="cuda")
idx = torch.randint(0, N, (B, N, k), device="cuda", dtype=torch.int32)
s1, t1, s2, t2 = (torch.randn(C, device="cuda") for _ in range(4))
w2 = torch.randn(C, C, device="cuda") / C ** 0.5
x = torch.empty(M, 512, device="cuda")
for mode in ("fp32", "tf32"):
    ops.set_precision(mode)
    ms = t(lambda: ops.edgeconv_dg(pq, 2 * C, pq[:, C:], 2 * C, idx, B, N, k, C, C, s1, t1, w2, s2, t2, ops.ACT_LEAKY, 0.01, x, 512, x[:, 128:], 512))
    print(f"edgeconv_dg 128x128 {mode}: {ms:.3f} ms ({2.0*M*k*C*C/ms/1e9:.1f} TFLOP/s)")
pq3 = torch.randn(M, 512, device="cuda")
ms = t(lambda: ops.edge_gather_ext(pq3, 512, pq3[:, 256:], 512, idx, B, N, k, 256, s1.repeat(2), t1.repeat(2), ops.ACT_LEAKY, 0.01, x[:, 256:], 512))
print(f"edge_gather_ext C=256: {ms:.3f} ms ({M*k*1024/ms/1e6:.0f} GB/s gathered)")
