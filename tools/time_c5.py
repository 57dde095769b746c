"""C5 stress shape: LPD-Net eval embedding on 16384-point clouds, k = 32 (BASELINE.json configs[4]); per-GPU share of the
256-cloud batch at 8 GPUs is 32.  usage: python tools/time_c5.py [B] [N] [k]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, synth
from lpdnet_b200.util.PointNetVlad import PointNetVlad

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
k = int(sys.argv[3]) if len(sys.argv) > 3 else 32
ops.set_precision("tf32")
model = PointNetVlad(num_points=N, featnet="lpdnet", emb_dims=1024)
model.emb_nn.k = k
model.load_state_dict(synth.synthetic_state_dict(model))
model = model.cuda().eval()
x = synth.clouds(B, N).cuda()
with torch.no_grad():
    for _ in range(2):
        out = model(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        out = model(x)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    ops.profile(True)
    model(x)
    rec = ops.profile(False)
    torch.cuda.synchronize()
print(f"C5 B={B} N={N} k={k}: {ms:.2f} ms/step = {B / ms * 1e3:.0f} submaps/s ({B * N / ms / 1e3:.1f} M points/s); peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
for label, a, b in sorted(rec, key=lambda r: -r[1].elapsed_time(r[2]))[:8]:
    print(f"   {label:40s} {a.elapsed_time(b):8.3f} ms")
