"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference modules (from /root/reference in the authoring container, or from
the staged copy oracle/_ref/ on the GPU box, see oracle/make_ref.py) without touching them:

  * util/lpdnet_model.py hard-codes torch.device('cuda') (:123,:307,:338); on a CPU run the module-global `torch` of that
    module is replaced by a proxy whose .device(...) returns cpu and which forwards every other attribute;
  * evaluate.py cannot be imported (it imports util.initPara: argv parsing, NVML init, directory creation), so get_recall
    (:162-206) is extracted from its AST and exec'd with np, KDTree and recall_num = 25.
"""
from __future__ import annotations

import ast
import os
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent


def reference_root() -> Path | None:
    """the reference tree to import from: $LPD_REFERENCE, /root/reference, else the staged oracle/_ref/; None if absent"""
    for cand in (os.environ.get("LPD_REFERENCE"), "/root/reference", HERE / "_ref"):
        if cand and (Path(cand) / "util" / "PointNetVlad.py").exists():
            return Path(cand)
    return None


def import_reference(root: Path | None = None, cpu: bool = True):
    """-> (util.lpdnet_model, util.PointNetVlad, loss.pointnetvlad_loss) of the reference"""
    import torch
    root = root or reference_root()
    if root is None:
        raise FileNotFoundError("no reference tree: neither /root/reference nor oracle/_ref/ (python -m oracle.make_ref) exists")
    sys.dont_write_bytecode = True
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    import util.lpdnet_model as L  # noqa

    if cpu:
        class _Proxy(types.ModuleType):
            def __getattr__(self, n):
                return getattr(torch, n)

            def device(self, *a, **k):
                return torch.device("cpu")

        L.torch = _Proxy("torch_proxy")
    import util.PointNetVlad as PNV  # noqa
    import loss.pointnetvlad_loss as RL  # noqa
    return L, PNV, RL


def extract_get_recall(root: Path | None = None):
    import numpy as np
    from sklearn.neighbors import KDTree
    root = root or reference_root()
    tree = ast.parse((Path(root) / "evaluate.py").read_text())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_recall")
    ns = {"np": np, "KDTree": KDTree, "recall_num": 25}
    exec(compile(ast.Module([node], []), "evaluate.py", "exec"), ns)
    return ns["get_recall"]
