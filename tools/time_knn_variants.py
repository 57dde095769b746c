"""Times the three formulations of the feature-space kNN on REAL conv2 features of the C2 workload (and optionally the C5
shape), checks them against the CUDA-core kernel bit for bit and reports the tiles that needed the exact fallback.
usage: python tools/time_knn_variants.py [B] [N] [k] [variants, e.g. 12]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, synth
from lpdnet_b200.util.PointNetVlad import PointNetVlad


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
variants = [int(c) for c in sys.argv[4]] if len(sys.argv) > 4 else [0, 1, 2, 3]
model = PointNetVlad(num_points=N, featnet="lpdnet", emb_dims=1024)
model.load_state_dict(synth.synthetic_state_dict(model))
model = model.cuda().eval()
x = synth.clouds(B, N).cuda()
emb = model.emb_nn
p = emb._prep.get(emb, emb._build)
with torch.no_grad():
    h, xyz, _, _ = emb._front(x, p, "LPDNet", True)
feat = h.view(B, N, 64).contiguous()
print(f"B={B} N={N} k={k}  feature norms: mean {feat.norm(dim=2).mean().item():.3f} max {feat.norm(dim=2).max().item():.3f}")
ops.KNN_TENSOR_CORES = False
ref = ops.knn(feat, k)
if B * N <= 64 * 4096 and len(variants) >= 3:
    print(f"knn_simt            : {t(lambda: ops.knn(feat, k), 3):8.3f} ms")
ops.KNN_TENSOR_CORES = True
for v in variants:
    ops.knn_tc_variant(v)
    diag = {}
    idx = ops.knn(feat, k, diag=diag)
    same = bool(torch.equal(idx, ref))
    ms = t(lambda: ops.knn(feat, k))
    print(f"knn_tc variant {v}    : {ms:8.3f} ms   bit-exact vs simt: {same}   mismatching rows: {int((idx != ref).any(dim=2).sum())}"
          f"   flagged tiles: {diag['flagged_tiles']} of {diag['tiles']}")
