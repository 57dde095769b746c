"""C4 timing on the GPU: all 506 run pairs in one pass (stacked 23 x 956-row databases, 3,036 queries) and the one-big-database
variant (21,988 rows), per-kernel CUDA-event breakdown.  usage: python tools/time_c4.py"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from lpdnet_b200 import evaluate, ops, synth  # noqa: E402

DB, Q, SETS = synth.descriptor_database()
dev = torch.device("cuda:0")
DBd = [torch.from_numpy(d).to(dev) for d in DB]
Qd = [torch.from_numpy(q).to(dev) for q in Q]
truth = evaluate.prepare_truth(SETS)
t0 = time.perf_counter()
evaluate.prepare_truth(SETS)
print(f"prepare_truth (host, once per evaluation set): {1e3 * (time.perf_counter() - t0):.1f} ms")
for _ in range(3):
    res = evaluate.recall_all_pairs(DBd, Qd, SETS, truth)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    res = evaluate.recall_all_pairs(DBd, Qd, SETS, truth)
e.record()
torch.cuda.synchronize()
print(f"recall_all_pairs (506 pairs, 66,792 query-database searches): {s.elapsed_time(e) / 10:.3f} ms per pass (device + host glue)")
ops.profile(True)
evaluate.recall_all_pairs(DBd, Qd, SETS, truth)
rec = ops.profile(False)
torch.cuda.synchronize()
for label, a, b in rec:
    print(f"   {label:40s} {a.elapsed_time(b):.3f} ms")
db = torch.cat(DBd, 0)
q = torch.cat(Qd, 0)
for fn, name in ((ops.retrieval_search, "tensor-core filter + refine"), (ops.retrieval_topk, "fp64 brute force")):
    for _ in range(2):
        fn(db, q, 25)
    torch.cuda.synchronize()
    s.record()
    for _ in range(5):
        i, d = fn(db, q, 25)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"big database 3,036 x 21,988 x 256, {name}: {ms:.3f} ms = {2 * 3036 * 21988 * 256 / ms / 1e9:.1f} TFLOP/s algorithmic")
a, b = ops.retrieval_search(db, q, 25), ops.retrieval_topk(db, q, 25)
print("bit-identical:", bool(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])))
