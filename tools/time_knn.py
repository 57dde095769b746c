"""Times lpd_knn on the C2 workload shapes (B=64, N=4096, k=20) for C=3 and C=64."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, synth

def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

B, N, k = 64, 4096, 20
xyz = synth.clouds(B, N)[:, 0].contiguous().cuda()
g = torch.Generator().manual_seed(1)
feat = torch.nn.functional.leaky_relu(torch.randn(B, N, 64, generator=g), 0.01).cuda()
print(f"knn C=3  : {t(lambda: ops.knn(xyz, k)):.3f} ms")
print(f"knn C=64 : {t(lambda: ops.knn(feat, k)):.3f} ms")
print(f"knn C=3 k=32 N=16384 B=8: {t(lambda: ops.knn(synth.clouds(8, 16384)[:, 0].contiguous().cuda(), 32), 2):.3f} ms (incl. H2D)")
