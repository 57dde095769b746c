"""One feature-space kNN call per refine mode on C2 features, for ncu captures: python tools/prof_knn.py [B] [N] [k]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, synth
from lpdnet_b200.util.PointNetVlad import PointNetVlad
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
model = PointNetVlad(num_points=N, featnet="lpdnet", emb_dims=1024)
model.load_state_dict(synth.synthetic_state_dict(model))
model = model.cuda().eval()
emb = model.emb_nn
p = emb._prep.get(emb, emb._build)
with torch.no_grad():
    h, xyz, _, _ = emb._front(synth.clouds(B, N).cuda(), p, "LPDNet", True)
feat = h.view(B, N, 64).contiguous()
for mode in (1, 0):
    ops.knn_tc_refine(mode)
    for _ in range(2):
        ops.knn(feat, k)
torch.cuda.synchronize()
