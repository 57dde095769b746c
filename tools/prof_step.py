"""One eval step of the C2 workload (B=64) for ncu captures: python tools/prof_step.py [precision] [B]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, synth
from lpdnet_b200.util.PointNetVlad import PointNetVlad
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ops.set_precision(prec)
model = PointNetVlad(num_points=4096, featnet="lpdnet", emb_dims=1024)
model.load_state_dict(synth.synthetic_state_dict(model))
model = model.cuda().eval()
x = synth.clouds(B, 4096).cuda()
with torch.no_grad():
    for _ in range(2):
        out = model(x)
torch.cuda.synchronize()
print(out.shape)
