"""Times the feature-space kNN on REAL conv2 features of the C2 workload (low intrinsic dimension, unlike i.i.d. noise),
and reports how many 64-row tiles needed the exact fallback."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import ops, synth, _lib
from lpdnet_b200.util.PointNetVlad import PointNetVlad

def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

B, N, k = 64, 4096, 20
model = PointNetVlad(num_points=N, featnet="lpdnet", emb_dims=1024)
model.load_state_dict(synth.synthetic_state_dict(model))
model = model.cuda().eval()
x = synth.clouds(B, N).cuda()
emb = model.emb_nn
p = emb._prep.get(emb, emb._build)
with torch.no_grad():
    h, xyz, _, _ = emb._front(x, p, "LPDNet")
feat = h.view(B, N, 64)
print("feature norms: mean %.3f max %.3f" % (feat.norm(dim=2).mean().item(), feat.norm(dim=2).max().item()))
lib = _lib.load()
idx = torch.empty(B, N, k, device="cuda", dtype=torch.int32)
nb = lib.lpd_knn_workspace_bytes(B, N, 64, k)
ws = torch.zeros((nb + 15) // 16 * 4, device="cuda")
st = torch.cuda.current_stream().cuda_stream
lib.lpd_knn_tc(feat.data_ptr(), B, N, 64, k, idx.data_ptr(), 0, ws.data_ptr(), ws.numel() * 4, st)
torch.cuda.synchronize()
npad = (N + 127) // 128 * 128
flags = ws.view(torch.int32)[B * npad + B: B * npad + B + B * ((N + 63) // 64)]
print("flagged 64-row tiles: %d of %d" % (int((flags != 0).sum()), flags.numel()))
print(f"knn_tc  C=64 real features: {t(lambda: ops.knn(feat, k)):.3f} ms")
ops.KNN_TENSOR_CORES = False
print(f"knn_simt C=64 real features: {t(lambda: ops.knn(feat, k)):.3f} ms")
