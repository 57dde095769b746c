import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, synth
for (B, N, k) in ((64, 4096, 20), (32, 16384, 32)):
    x = synth.clouds(B, N)[:, 0].contiguous().cuda()
    for grid in (False, True):
        ops.KNN_GRID = grid
        for _ in range(3): idx = ops.knn(x, k)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): idx = ops.knn(x, k)
        e.record(); torch.cuda.synchronize()
        print(f"B={B} N={N} k={k} grid={grid}: {s.elapsed_time(e)/10:.3f} ms")
        if grid: same = torch.equal(idx, ref)
        else: ref = idx
    print("  identical:", same)
