import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from lpdnet_b200 import ops
torch.manual_seed(0)
for (M,N,K,b) in ((128,64,32,1),(128,64,64,1),(256,128,256,2)):
    A = torch.randn(b*K, M, device="cuda"); B = torch.randn(b*K, N, device="cuda")
    out = torch.full((b, M, N), 7.0, device="cuda")
    ops.gemm_tf32_tn(A, B, M=M, N=N, K=K, lda=M, ldb=N, batch=b, out=out)
    torch.cuda.synchronize()
    ref = torch.einsum("zkm,zkn->zmn", A.view(b,K,M).double(), B.view(b,K,N).double())
    print(M,N,K,b, "max err", float((out.double()-ref).abs().max()), "ref max", float(ref.abs().max()), "out[0,:2,:4]", out[0,:2,:4].tolist(), "ref", ref[0,:2,:4].tolist())
