"""Generates tests/golden/host_pipeline.npz from the UNMODIFIED reference functions rotate_point_cloud, jitter_point_cloud and
get_query_tuple (loading_pointclouds.py:50-142), pulled out by AST extraction because the module itself cannot be imported
(it drags in util.initPara: argparse, NVML, log files).  Run in the authoring container only (needs /root/reference):
    python oracle/gen_golden_host.py
The submap reader is replaced by a stub that encodes the file name in the array, so the fixture records WHICH submaps a seeded
call selects, in which order."""
import ast
import copy
import random
from pathlib import Path
from time import time

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "host_pipeline.npz"


def stub_load_pc_file(filename, *_):
    return np.full((8, 3), float(filename), dtype=np.float64)


def stub_load_pc_files(filenames, *_):
    return np.array([stub_load_pc_file(f) for f in filenames])


def extract(names):
    tree = ast.parse((REF / "loading_pointclouds.py").read_text())
    nodes = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {"np": np, "random": random, "time": time, "load_pc_file": stub_load_pc_file, "load_pc_files": stub_load_pc_files}
    exec(compile(ast.Module(nodes, []), "loading_pointclouds.py", "exec"), ns)
    return ns


def query_dict(n=40, seed=3):
    r = np.random.default_rng(seed)
    d = {}
    for i in range(n):
        pos = [j for j in range(n) if j != i and abs(j - i) <= 2]
        neg = [j for j in range(n) if abs(j - i) > 5]
        r.shuffle(neg)
        d[i] = {"query": str(1000 + i), "positives": pos, "negatives": neg}
    return d


def main():
    ns = extract({"rotate_point_cloud", "jitter_point_cloud", "get_query_tuple"})
    out = {}
    base = np.random.default_rng(11).uniform(-1, 1, (3, 50, 3))
    out["clouds"] = base
    np.random.seed(7)
    out["rotated"] = ns["rotate_point_cloud"](base)
    out["jittered"] = ns["jitter_point_cloud"](base)
    for case, (hard, other) in enumerate([([], False), ([], True), ([20, 31], True)]):
        qd = query_dict()
        random.seed(100 + case)
        res = ns["get_query_tuple"](copy.deepcopy(qd[7]), 2, 6, qd, hard_neg=hard, other_neg=other)
        for name, arr in zip(("q", "pos", "neg", "other"), res):
            out[f"tuple{case}_{name}"] = np.asarray(arr)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
