"""Times lpd_gemm_f16 (fp16 in / fp16 out: the CTA-pair kernel for N % 256 == 0) on the wide layers of the C2 step against cuBLAS fp16
(library call: yardstick only).  usage: python tools/time_gemm_f16.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops


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


for M, N, K in ((262144, 1024, 512), (262144, 512, 128)):
    A = torch.randn(M, K, device="cuda").half()
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    sc, sh = torch.rand(N, device="cuda") + 0.5, torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    a = t(lambda: ops.gemm_f16(A, W, M=M, N=N, K=K, out=out, scale=sc, shift=sh, act=ops.ACT_LEAKY, slope=0.01))
    c = t(lambda: torch.matmul(A, W.t(), out=out))
    fl = 2.0 * M * N * K
    print(f"{M}x{N}x{K}: lpd_gemm_f16 {a:.3f} ms ({fl / a / 1e9:.0f} TF/s)   cuBLAS fp16 (no epilogue) {c:.3f} ms ({fl / c / 1e9:.0f} TF/s)")
