"""TEST INFRASTRUCTURE — recipe that stages the UNMODIFIED reference modules of the hot path under oracle/_ref/.

The reference is pure Python (no C/C++ to compile), so "building" it means copying the five files the path consists of,
byte for byte, from where they lie under /root/reference into oracle/_ref/ (git-ignored, NOT gpurun-ignored: like a built
.so it travels to the GPU box, where /root/reference does not exist).  Nothing under oracle/_ref/ is ever committed,
edited or imported by the product; bench.py's `--impl reference` arm and `cpu_baseline` leg import it through
oracle/ref_loader.py (kind == "reference"), and fall back to the numpy/C port (kind == "port") when it is absent.

    python -m oracle.make_ref            # also run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("LPD_REFERENCE", "/root/reference"))
OUT = HERE / "_ref"

# file -> why it is needed
FILES = {
    "util/__init__.py": "package marker",
    "util/lpdnet_model.py": "LPDNet, LPDNetOrign, TranformNet, knn, get_graph_feature[_Origin]",
    "util/PointNetVlad.py": "PointNetVlad, NetVLADLoupe, GatingContext, STN3d, PointNetfeat",
    "util/gpu_mem_track.py": "imported by util/lpdnet_model.py at import time (MemTracker construction only)",
    "loss/pointnetvlad_loss.py": "best_pos_distance, triplet_loss, quadruplet_loss",
    "evaluate.py": "get_recall (:162-206) — never imported (import-time side effects); extracted from its AST",
}


def make(verbose: bool = True) -> bool:
    if not REF.exists():
        if verbose:
            print(f"[make_ref] {REF} not present: keeping whatever oracle/_ref/ already holds")
        return OUT.exists()
    manifest = {}
    for rel in FILES:
        src, dst = REF / rel, OUT / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(dst.read_bytes()).hexdigest()
    (OUT / "loss" / "__init__.py").touch()
    (OUT / "MANIFEST.json").write_text(json.dumps({"source": str(REF), "sha256": manifest}, indent=1))
    if verbose:
        print(f"[make_ref] staged {len(FILES)} reference files under {OUT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
