"""Times the fused conv1+conv2 input kernel against the two strict-fp32 GEMMs it replaces (C2 shape)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops
M = 64 * 4096
x = torch.randn(M, 3, device="cuda")
w1, w2 = torch.randn(64, 3, device="cuda"), torch.randn(64, 64, device="cuda") / 8
s1, t1, s2, t2 = (torch.randn(64, device="cuda") for _ in range(4))


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def two():
    h = ops.gemm(x, w1, M=M, N=64, K=3, lda=3, scale=s1, shift=t1, act=ops.ACT_LEAKY, slope=0.01)
    return ops.gemm(h, w2, M=M, N=64, K=64, scale=s2, shift=t2, act=ops.ACT_LEAKY, slope=0.01)


a = ops.pointwise_mlp2(x, 3, M, w1, s1, t1, w2, s2, t2, ops.ACT_LEAKY, 0.01)
print("max |fused - two layers|:", float((a - two()).abs().max()))
print(f"fused {t(lambda: ops.pointwise_mlp2(x, 3, M, w1, s1, t1, w2, s2, t2, ops.ACT_LEAKY, 0.01)):.4f} ms   two GEMMs {t(two):.4f} ms")
