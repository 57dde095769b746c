import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent/"tests"))
import numpy as np, torch
from lpdnet_b200 import ops, synth
from test_train_gpu import build_train, run_step, subsample
for name, N, Bq in (("c3_train_step_n256", 256, 1), ("c3_train_step_n512_b2", 512, 2)):
    g = np.load(f"tests/golden/{name}.npz")
    ops.set_precision("fp32")
    model = build_train(N)
    out, loss = run_step(model, synth.clouds(Bq*22, N), Bq)
    print(name, "out err", np.abs(out.detach().cpu().numpy()-g["out"]).max(), "loss", float(loss), float(g["loss"]))
    for key, p in model.named_parameters():
        ref = g["grad."+key]; got = subsample(p.grad)
        e = np.abs(got-ref).max()/max(np.abs(ref).max(),1e-12)
        gn = float(p.grad.double().norm()); rn=float(g["gnorm."+key])
        flag = " <<<" if e>5e-4 or abs(gn-rn)>2e-4*rn else ""
        print(f"  {key:45s} maxrel {e:.2e} normrel {abs(gn-rn)/max(rn,1e-12):.2e}{flag}")
        if e > 2e-3:
            d = np.abs(got-ref); i = np.argsort(-d)[:5]; print("     worst idx", i, got[i], ref[i])
