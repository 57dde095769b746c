"""Times the Cartesian (grid) kNN on the C2 / C5 shapes and checks it against the brute-force CUDA-core kernel.
usage: python tools/time_knn_xyz.py [B] [N] [k]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
x = synth.clouds(B, N)[:, 0].contiguous().cuda()


def t(fn, n=10):
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


ops.KNN_GRID = False
ref = ops.knn(x, k)
brute = t(lambda: ops.knn(x, k), 3)
ops.KNN_GRID = True
got = ops.knn(x, k)
print(f"B={B} N={N} k={k}: grid == brute force: {bool(torch.equal(got, ref))}  mismatching rows {int((got != ref).any(dim=2).sum())}")
print(f"brute force {brute:.3f} ms   grid {t(lambda: ops.knn(x, k)):.3f} ms")
