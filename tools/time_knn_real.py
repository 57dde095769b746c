"""Times the feature-space kNN on REAL conv2 features of the C2 workload (low intrinsic dimension, unlike i.i.d. noise),
checks the tensor-core path against the CUDA-core kernel bit for bit and reports how many 64-row tiles needed the fallback."""
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

B, N = 64, 4096
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
model = PointNetVlad(num_points=N, featnet="lpdnet", emb_dims=1024)
model.load_state_dict(synth.synthetic_state_dict(model))
model = model.cuda().eval()
x = synth.clouds(B, N).cuda()
emb = model.emb_nn
p = emb._prep.get(emb, emb._build)
with torch.no_grad():
    h, xyz, _, _ = emb._front(x, p, "LPDNet", True)
feat = h.view(B, N, 64)
print("feature norms: mean %.3f max %.3f" % (feat.norm(dim=2).mean().item(), feat.norm(dim=2).max().item()))
lib = _lib.load()
idx = torch.empty(B, N, k, device="cuda", dtype=torch.int32)
nb = lib.lpd_knn_workspace_bytes(B, N, 64, k)
ws = torch.zeros((nb + 3) // 4, device="cuda")
st = torch.cuda.current_stream().cuda_stream
rc = lib.lpd_knn_tc(feat.data_ptr(), B, N, 64, k, idx.data_ptr(), 0, ws.data_ptr(), ws.numel() * 4, st)
torch.cuda.synchronize()
print("rc", rc)
npad = (N + 127) // 128 * 128
off_flags = ((B * npad * 4 + 255) // 256 * 256 + B * 4) // 4
flags = ws.view(torch.int32)[off_flags: off_flags + B * ((N + 63) // 64)]
print("flagged 64-row tiles: %d of %d" % (int((flags != 0).sum()), flags.numel()))
ops.KNN_TENSOR_CORES = False
ref = ops.knn(feat, k)
print("tc == simt (bit-exact):", bool(torch.equal(idx, ref)), " mismatching rows:", int((idx != ref).any(dim=2).sum()))
print(f"knn_simt C=64 real features: {t(lambda: ops.knn(feat, k)):.3f} ms")
ops.KNN_TENSOR_CORES = True
print(f"knn_tc   C=64 real features: {t(lambda: ops.knn(feat, k)):.3f} ms")
ops.profile(True)
for _ in range(3): ops.knn(feat, k)
rec = ops.profile(False)
torch.cuda.synchronize()
print("per-call ms:", [round(a.elapsed_time(b), 3) for _, a, b in rec])
