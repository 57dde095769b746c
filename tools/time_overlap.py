"""Does the Cartesian kNN overlap with the feature-space kNN when launched on a low-priority side stream?
usage: python tools/time_overlap.py [B]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, synth
from lpdnet_b200.util.PointNetVlad import PointNetVlad

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N, k = 4096, 20
model = PointNetVlad(num_points=N, featnet="lpdnet", emb_dims=1024)
model.load_state_dict(synth.synthetic_state_dict(model))
model = model.cuda().eval()
x = synth.clouds(B, N).cuda()
emb = model.emb_nn
p = emb._prep.get(emb, emb._build)
with torch.no_grad():
    h, xyz, _, _ = emb._front(x, p, "LPDNet", True)
feat = h.view(B, N, 64).contiguous()
xyz = xyz.contiguous()
lo, hi = torch.cuda.Stream.priority_range()
main = torch.cuda.Stream(priority=hi)
side = torch.cuda.Stream(priority=lo)
side_hi = torch.cuda.Stream(priority=hi)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(main)
    for _ in range(n):
        fn()
    e.record(main)
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def serial():
    with torch.cuda.stream(main):
        ops.knn(feat, k)
        ops.knn(xyz, k)


def overlapped(side_stream, side_first):
    def f():
        ev0, ev1 = torch.cuda.Event(), torch.cuda.Event()
        ev0.record(main)
        side_stream.wait_event(ev0)
        if side_first:
            with torch.cuda.stream(side_stream):
                ops.knn(xyz, k)
                ev1.record(side_stream)
            with torch.cuda.stream(main):
                ops.knn(feat, k)
        else:
            with torch.cuda.stream(main):
                ops.knn(feat, k)
            with torch.cuda.stream(side_stream):
                ops.knn(xyz, k)
                ev1.record(side_stream)
        main.wait_event(ev1)
    return f


print(f"serial                                   {timeit(serial):.3f} ms")
print(f"side stream low priority, side first     {timeit(overlapped(side, True)):.3f} ms")
print(f"side stream low priority, main first     {timeit(overlapped(side, False)):.3f} ms")
print(f"side stream same priority, side first    {timeit(overlapped(side_hi, True)):.3f} ms")
print(f"side stream same priority, main first    {timeit(overlapped(side_hi, False)):.3f} ms")
