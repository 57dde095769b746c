import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent/"tests"))
import numpy as np, torch, traceback
from lpdnet_b200 import ops, synth
from lpdnet_b200.util import PointNetVlad as PNV
from test_train_gpu import run_step, subsample
cases = [("train_step_lpdnetorigin_n256", dict(featnet="lpdnetorigin"), 1234),
    ("train_step_pointnet_n256", dict(featnet="pointnet"), 4),
    ("train_step_pointnet_ft_n256", dict(featnet="pointnet", feature_transform=True), 1234),
    ("train_step_lpdnet_tnets_n256", dict(featnet="lpdnet", feature_transform=True, xyz_trans=True), 1234),
    ("train_step_lpdnetorigin_tnets_n256", dict(featnet="lpdnetorigin", feature_transform=True, xyz_trans=True), 1234)]
import os
ops.SPATIAL_ORDER = os.environ.get("SPATIAL","1") == "1"
for name, kw, seed in cases[1:3]:
    g = np.load(f"tests/golden/{name}.npz")
    ops.set_precision("fp32")
    model = PNV.PointNetVlad(num_points=256, emb_dims=1024, **kw)
    model.load_state_dict(synth.synthetic_state_dict(model)); model = model.cuda().train()
    try:
        out, loss = run_step(model, synth.clouds(22, 256, seed=seed), 1)
    except Exception:
        print(name, "FAILED"); traceback.print_exc(); continue
    o = out.detach().cpu().numpy()
    print(name, "out err32", np.abs(o-g["out"]).max(), "err64", np.abs(o-g["out64"]).max(), "noise", np.abs(g["out"]-g["out64"]).max(), "loss", float(loss.detach()), float(g["loss"]), float(g["loss64"]))
    for key, p in model.named_parameters():
        if "grad."+key not in g.files:
            print(f"  {key:45s} ref has no grad; mine {'None' if p.grad is None else 'SET'}"); continue
        if p.grad is None: print(f"  {key:45s} MISSING GRAD"); continue
        ref, ref64 = g["grad."+key], g["grad64."+key]; got = subsample(p.grad)
        sc = max(np.abs(ref64).max(),1e-12)
        e32, e64, nz = np.abs(got-ref).max()/sc, np.abs(got-ref64).max()/sc, np.abs(ref-ref64).max()/sc
        gn = float(p.grad.double().norm()); rn=float(g["gnorm64."+key])
        flag = " <<<" if e64 > max(5e-4, 2*nz) else ""
        print(f"  {key:45s} e32 {e32:.1e} e64 {e64:.1e} noise {nz:.1e} normrel {abs(gn-rn)/max(rn,1e-12):.1e}{flag}")
