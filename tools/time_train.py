"""One C3 training step at full size (44 clouds x 4096 pts) with a per-kernel breakdown: python tools/time_train.py [precision] [B]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, optim, synth
from lpdnet_b200.loss import pointnetvlad_loss as L
from lpdnet_b200.util.PointNetVlad import PointNetVlad
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
Bq = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ops.set_precision(prec)
model = PointNetVlad(num_points=4096, featnet="lpdnet", emb_dims=1024)
model.load_state_dict(synth.synthetic_state_dict(model))
model = model.cuda().train()
opt = optim.Adam(model.parameters(), lr=1e-3)
x = synth.clouds(Bq * 22, 4096).cuda()

def step():
    opt.zero_grad()
    out = model(x)
    q, pos, neg, other = torch.split(out.view(Bq, -1, 256), [1, 2, 18, 1], dim=1)
    loss = L.quadruplet_loss(q, pos, neg, other, 0.5, 0.2, use_min=True, lazy=True)
    loss.backward()
    opt.step()
    return loss

for _ in range(3):
    loss = step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
n = 5
for _ in range(n):
    loss = step()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / n
print(f"precision {prec}: {ms:.2f} ms/step, {Bq*22/ms*1e3:.1f} submaps/s, loss {float(loss.detach()):.4f}, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
ops.profile(True)
step()
rec = ops.profile(False)
torch.cuda.synchronize()
tot = {}
for label, a, b in rec:
    t = tot.setdefault(label, [0.0, 0]); t[0] += a.elapsed_time(b); t[1] += 1
allms = sum(v[0] for v in tot.values())
print(f"sum of kernels {allms:.2f} ms")
for k_, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"  {k_:45s} {v[0]:8.3f} ms  x{v[1]}")
