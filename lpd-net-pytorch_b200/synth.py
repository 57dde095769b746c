"""Seeded synthetic inputs and weights shared by the tests, the golden-vector generator and bench.py.

The reference ships neither data nor its checkpoint (pretrained/lpdnet.ckpt is a missing blob), so
every parity run uses:
  * clouds: uniform [-1, 1]^3, seed 1234 (the reference's default --seed, util/initPara.py:86; its own
    smoke input is torch.rand(44, 1, 4096, 3), util/PointNetVlad.py:275);
  * weights: a deterministic fill keyed on the state_dict KEY NAME, so the reference model and this
    package's model receive bit-identical tensors through load_state_dict regardless of constructor
    RNG order.  BatchNorm statistics and affine terms are randomised (negative scales included) so a
    folded-BN path is really exercised.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def clouds(B: int, N: int = 4096, seed: int = 1234, dims: int = 3) -> torch.Tensor:
    """[B, 1, N, dims] float32 on the CPU, uniform in [-1, 1]."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 1, N, dims, generator=g) * 2 - 1


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(key.encode())])


def fill_state_dict(shapes: dict, seed: int = 4321) -> dict:
    """shapes: {key: torch.Size / tuple}  ->  {key: torch.Tensor} (float32 / int64), deterministic per key."""
    bn_prefixes = {k[: -len("running_mean")] for k in shapes if k.endswith("running_mean")}
    out = {}
    for key, shape in shapes.items():
        shape = tuple(shape)
        r = _rng(seed, key)
        prefix, _, leaf = key.rpartition(".")
        prefix = prefix + "." if prefix else ""
        if leaf == "num_batches_tracked":
            out[key] = torch.zeros(shape, dtype=torch.int64)
            continue
        if prefix in bn_prefixes:
            if leaf == "running_mean":
                a = r.normal(0.0, 0.1, shape)
            elif leaf == "running_var":
                a = r.uniform(0.5, 2.0, shape)
            elif leaf == "weight":
                a = r.normal(0.0, 1.0, shape)
            else:  # bias
                a = r.normal(0.0, 0.1, shape)
        elif leaf in ("cluster_weights", "cluster_weights2", "hidden1_weights"):
            feat = shape[-2] if leaf != "hidden1_weights" else shape[0] // 64
            a = r.normal(0.0, 1.0, shape) / np.sqrt(feat)
        elif leaf == "gating_weights":
            a = r.normal(0.0, 1.0, shape) / np.sqrt(shape[0])
        elif len(shape) >= 2:  # conv / linear weight
            fan_in = int(np.prod(shape[1:]))
            a = r.normal(0.0, 1.0, shape) / np.sqrt(fan_in)
        else:  # conv / linear bias, cluster_biases, gating_biases
            a = r.normal(0.0, 0.1, shape)
        out[key] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return out


def synthetic_state_dict(model: torch.nn.Module, seed: int = 4321) -> dict:
    return fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed)


def descriptor_database(runs: int = 23, places: int = 956, queries_per_run: int = 132, dim: int = 256,
                        rho: float = 0.99, sigma: float = 0.6, seed: int = 1234):
    """Synthetic retrieval workload of SURVEY §8(d) C4: a trajectory of correlated place descriptors,
    `runs` noisy traversals (database sets) and per-run query subsets with ground truth {p-1, p, p+1}.

    Returns (DATABASE_VECTORS, QUERY_VECTORS, QUERY_SETS) in the structures evaluate.get_recall expects:
    lists of float32 arrays and QUERY_SETS[n][i][m] = list of true database indices in run m.
    """
    r = np.random.default_rng(seed)
    z = np.empty((places, dim))
    z[0] = r.standard_normal(dim)
    z[0] /= np.linalg.norm(z[0])
    for p in range(1, places):
        v = rho * z[p - 1] + np.sqrt(1 - rho * rho) * r.standard_normal(dim) / np.sqrt(dim)
        z[p] = v / np.linalg.norm(v)
    db, qs, sets = [], [], []
    q_places = np.sort(r.choice(places, size=queries_per_run, replace=False))
    for _ in range(runs):
        noisy = z + sigma * r.standard_normal((places, dim)) / np.sqrt(dim)
        noisy /= np.linalg.norm(noisy, axis=1, keepdims=True)
        db.append(noisy.astype(np.float32))
    for n in range(runs):
        qs.append(db[n][q_places].copy())
        per_q = []
        for i, p in enumerate(q_places):
            truth = {}
            for m in range(runs):
                # leave some truths empty to exercise the skip at evaluate.py:181-182
                if (i + 3 * m + n) % 17 == 0:
                    truth[m] = []
                else:
                    truth[m] = [int(t) for t in (p - 1, p, p + 1) if 0 <= t < places]
            per_q.append(truth)
        sets.append(per_q)
    return db, qs, sets
