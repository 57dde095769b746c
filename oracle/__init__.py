"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A CPU restatement of the reference's algorithm for the hot path (qiaozhijian/LPD-Net-Pytorch):
  knn_canonical.c   canonical kNN (util/lpdnet_model.py:317-326) and fp64 brute-force retrieval
                    (evaluate.py:168,186-187)
  model_numpy.py    LPDNet / LPDNetOrign / PointNetfeat / STN3d / TranformNet / NetVLADLoupe /
                    GatingContext forward, as written (materialised edge tensors), in numpy
  loss_numpy.py     best_pos_distance / triplet_loss / quadruplet_loss (+ analytic gradients)
  recall_numpy.py   get_recall (evaluate.py:162-206)
  gen_golden.py     runs the REAL reference (imported from /root/reference, this container only) on the
                    seeded synthetic inputs and writes tests/golden/*.npz

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker.  The product (lpd-net-pytorch_b200/) never does.

Pinning: the reference has no tests and no golden vectors of its own, so the oracle is
pinned against outputs of the reference itself, generated here by gen_golden.py and committed under
tests/golden/ (see tests/test_oracle_vs_golden.py).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile the C part of the oracle (gcc, a second or two)."""
    if force or not _SO.exists() or _SO.stat().st_mtime < (_HERE / "knn_canonical.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(_SO))
        lib.lpd_oracle_knn.restype = None
        lib.lpd_oracle_knn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.lpd_oracle_retrieval.restype = None
        lib.lpd_oracle_retrieval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def knn_canonical(x_pm: np.ndarray, k: int, threads: int = 8, return_pd: bool = False):
    """x_pm [B, N, C] float32 point-major -> idx [B, N, k] int32 (canonical order, see knn_canonical.c)."""
    lib = _load()
    x_pm = np.ascontiguousarray(x_pm, dtype=np.float32)
    B, N, Cc = x_pm.shape
    idx = np.empty((B, N, k), dtype=np.int32)
    pd = np.empty((B, N, k), dtype=np.float32) if return_pd else None

    def one(b):
        lib.lpd_oracle_knn(x_pm[b].ctypes.data, 1, N, Cc, k, idx[b].ctypes.data,
                           pd[b].ctypes.data if return_pd else None)

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        list(ex.map(one, range(B)))
    return (idx, pd) if return_pd else idx


def retrieval_bruteforce(db: np.ndarray, q: np.ndarray, k: int, threads: int = 8):
    """fp64 brute-force k nearest rows of db for each row of q -> (idx int32 [Nq,k], sqdist float64 [Nq,k])."""
    lib = _load()
    db = np.ascontiguousarray(db, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    Nq, D = q.shape
    idx = np.empty((Nq, k), dtype=np.int32)
    dist = np.empty((Nq, k), dtype=np.float64)
    chunks = np.array_split(np.arange(Nq), max(1, min(threads, Nq)))

    def one(rows):
        if len(rows) == 0:
            return
        a, b = int(rows[0]), int(rows[-1]) + 1
        lib.lpd_oracle_retrieval(db.ctypes.data, db.shape[0], q[a:b].ctypes.data, b - a, D, k,
                                 idx[a:b].ctypes.data, dist[a:b].ctypes.data)

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        list(ex.map(one, chunks))
    return idx, dist
