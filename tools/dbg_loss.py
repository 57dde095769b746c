import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, optim, synth
from lpdnet_b200 import train_pointnetvlad as TP
from lpdnet_b200.util.PointNetVlad import PointNetVlad
import bench
ops.set_precision("tf32")
for lr in (1e-5, 1e-7):
    model = PointNetVlad(num_points=4096, featnet="lpdnet", emb_dims=1024)
    model.load_state_dict(synth.synthetic_state_dict(model)); model = model.cuda().train()
    opt = optim.Adam(model.parameters(), lr=lr)
    ls = []
    for i in range(8):
        b = tuple(t.cuda() for t in bench.synth_tuples(100 + i % 4))
        ls.append(float(TP.train_step(model, opt, *b)))
    print(lr, ls)
