"""How often does the tensor-core kNN filter hand a 64-row tile to the exact kernel on the C2 workload?  Runs the conv1/conv2 front
end on many seeded 64-cloud batches and counts flagged tiles.  usage: [LPD_KNN_CAP=40] python tools/knn_flag_stats.py [batches]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, synth
from lpdnet_b200.util.PointNetVlad import PointNetVlad
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 40
model = PointNetVlad(num_points=4096, featnet="lpdnet", emb_dims=1024)
model.load_state_dict(synth.synthetic_state_dict(model))
model = model.cuda().eval()
emb = model.emb_nn
p = emb._prep.get(emb, emb._build)
flagged, batches_hit, ms = 0, 0, []
for i in range(nb):
    x = synth.clouds(64, 4096, seed=5000 + i).cuda()
    with torch.no_grad():
        h, _, _, _ = emb._front(x, p, "LPDNet", True)
    feat = h.view(64, 4096, 64).contiguous()
    diag = {}
    ops.knn(feat, 20, diag=diag)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.knn(feat, 20); e.record(); torch.cuda.synchronize()
    ms.append(s.elapsed_time(e))
    flagged += diag["flagged_tiles"]
    batches_hit += diag["flagged_tiles"] > 0
print(f"{nb} batches of 64 x 4096: {flagged} flagged tiles in {batches_hit} batches; kNN ms min {min(ms):.3f} median {sorted(ms)[len(ms)//2]:.3f} max {max(ms):.3f}")
