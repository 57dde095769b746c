"""N > 1 host logic on CPU with the gloo backend, world size 2 (no GPU): database-sharded retrieval (local exact top-k,
all-gather, deterministic merge) against the unsharded oracle search; whole-tuple sharding and the gradient all-reduce
of the data-parallel training step (each rank's oracle gradients, summed, equal the sum computed in one process)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_local_topk(db, q, k, off):
    from oracle import retrieval_bruteforce
    idx, d = retrieval_bruteforce(db.numpy(), q.numpy(), k)
    return torch.from_numpy(idx.astype(np.int32) + off), torch.from_numpy(d.astype(np.float64))


def _oracle_merge(all_dst, all_idx):
    d = all_dst.permute(1, 0, 2).reshape(all_dst.shape[1], -1).numpy()
    i = all_idx.permute(1, 0, 2).reshape(all_idx.shape[1], -1).numpy().astype(np.int64)
    k = all_idx.shape[2]
    i_key = np.where(i < 0, np.iinfo(np.int64).max, i)
    order = np.lexsort((i_key, d), axis=1)[:, :k]                     # distance, then lower global index
    return torch.from_numpy(np.take_along_axis(i, order, 1).astype(np.int32)), torch.from_numpy(np.take_along_axis(d, order, 1))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lpdnet_b200 import parallel
    rng = np.random.default_rng(11)
    db = rng.standard_normal((203, 32)).astype(np.float32)
    db[17] = db[150]                                                  # an exact tie across the shard boundary
    q = rng.standard_normal((19, 32)).astype(np.float32)
    lo, hi = parallel.shard_range(len(db), world, rank)
    idx, dst = parallel.sharded_retrieval_topk(torch.from_numpy(db[lo:hi]), torch.from_numpy(q), 25, lo,
                                               local_topk=_oracle_local_topk, merge=_oracle_merge)
    # data-parallel gradient sum: every rank contributes its own "gradient" of its whole tuples
    tuples = (torch.arange(5 * 3.0).view(5, 3), torch.arange(5 * 2.0).view(5, 2))
    mine = parallel.shard_tuples(tuples, world, rank)
    g = torch.stack([t.sum() for t in mine])
    dist.all_reduce(g)
    cnt = parallel.allreduce_counters(torch.tensor([float(mine[0].shape[0])]))
    np.savez(Path(out_dir) / f"r{rank}.npz", idx=idx.numpy(), dst=dst.numpy(), g=g.numpy(), cnt=cnt.numpy(), lo=lo, hi=hi)
    dist.destroy_process_group()


def test_world2_sharded_retrieval_and_gradient_sum(tmp_path):
    from oracle import retrieval_bruteforce
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    rng = np.random.default_rng(11)
    db = rng.standard_normal((203, 32)).astype(np.float32)
    db[17] = db[150]
    q = rng.standard_normal((19, 32)).astype(np.float32)
    ref_idx, ref_d = retrieval_bruteforce(db, q, 25)
    for r in (r0, r1):
        assert np.array_equal(r["idx"], ref_idx.astype(np.int32))     # bit-exact indices, ties to the lower global index
        assert np.array_equal(r["dst"], ref_d.astype(np.float64))
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 102, 102, 203)
    assert np.array_equal(r0["g"], r1["g"]) and np.allclose(r0["g"], [np.arange(15.0).sum(), np.arange(10.0).sum()])
    assert float(r0["cnt"][0]) == 5.0                                 # 3 + 2 whole tuples


def test_shard_range_covers_everything_once():
    from lpdnet_b200 import parallel
    for n in (0, 1, 7, 44, 21988):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
