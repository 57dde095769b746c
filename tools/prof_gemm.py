"""One launch of each tensor-core GEMM shape, for ncu."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops
for M, N, K in ((262144, 1024, 512), (262144, 512, 128)):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
    out = torch.empty(M, N, device="cuda")
    for _ in range(2):
        ops.gemm_tf32(A, W, M=M, N=N, K=K, out=out, ldc=N)
    torch.cuda.synchronize()
