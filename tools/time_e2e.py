"""Where the end-to-end embedding path spends its time: pageable vs pinned input, graph replay vs eager (C2 shape)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lpdnet_b200 import evaluate, ops, synth
from lpdnet_b200.util.PointNetVlad import PointNetVlad

ops.set_precision("tf32")
model = PointNetVlad(num_points=4096, featnet="lpdnet", emb_dims=1024)
model.load_state_dict(synth.synthetic_state_dict(model))
model = model.cuda().eval()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 12
big = synth.clouds(64 * nb, 4096)[:, 0].contiguous()
pinned = big.pin_memory()
t0 = time.perf_counter(); tmp = torch.empty(64, 4096, 3).pin_memory(); t1 = time.perf_counter()
for _ in range(5):
    tmp.copy_(big[:64])
t2 = time.perf_counter()
print(f"pin_memory(3 MB) {1e3 * (t1 - t0):.2f} ms;  pageable -> pinned memcpy of one batch {1e3 * (t2 - t1) / 5:.3f} ms")
for name, src, g in (("pageable+graph", big, True), ("pinned+graph", pinned, True), ("pageable eager", big, False), ("pinned eager", pinned, False)):
    evaluate.get_latent_vectors(model, src[:192], batch_num=64, use_graph=g)
    torch.cuda.synchronize()
    t = time.perf_counter()
    evaluate.get_latent_vectors(model, src, batch_num=64, use_graph=g)
    dt = time.perf_counter() - t
    print(f"{name:16s}: {1e3 * dt / nb:.3f} ms per batch  ({64 * nb / dt:.0f} submaps/s)")
