"""Times lpd_gemm (fp32 FFMA) against lpd_gemm_tf32 (tcgen05) on the layer shapes of the C2 workload."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

for M, N, K in ((262144, 1024, 512), (262144, 512, 128), (262144, 256, 64), (262144, 64, 1024), (262144, 128, 128)):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
    out = torch.empty(M, N, device="cuda")
    a = t(lambda: ops.gemm(A, W, M=M, N=N, K=K, out=out, ldc=N))
    b = t(lambda: ops.gemm_tf32(A, W, M=M, N=N, K=K, out=out, ldc=N))
    torch.backends.cuda.matmul.allow_tf32 = True
    c = t(lambda: torch.matmul(A, W.t(), out=out))
    fl = 2.0 * M * N * K
    print(f"{M}x{N}x{K}: fp32 {a:.3f} ms ({fl/a/1e9:.1f} TF/s)  tf32 {b:.3f} ms ({fl/b/1e9:.1f} TF/s)  cublas-tf32 {c:.3f} ms ({fl/c/1e9:.1f} TF/s)")
