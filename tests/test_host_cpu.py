"""Host-side logic that needs no GPU: drop-in surface (constructor signatures, state_dict keys and shapes equal
to the reference's), the C-ABI library loads and exports every symbol the header declares, and the product
fails loudly instead of falling back when there is no CUDA tensor / no library."""
import inspect
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from lpdnet_b200 import _lib, synth
from lpdnet_b200.loss import pointnetvlad_loss as L
from lpdnet_b200.util import PointNetVlad as PNV
from lpdnet_b200.util import lpdnet_model as LM

ROOT = Path(__file__).resolve().parent.parent

MODEL_FIXTURES = [
    ("c1_pointnet_eval", dict(num_points=4096, featnet="pointnet", emb_dims=1024)),
    ("c2_lpdnet_eval", dict(num_points=4096, featnet="lpdnet", emb_dims=1024)),
    ("c2_lpdnet_tnets_eval", dict(num_points=1024, featnet="lpdnet", emb_dims=1024, feature_transform=True, xyz_trans=True)),
    ("c2_lpdnetorigin_eval", dict(num_points=4096, featnet="lpdnetorigin", emb_dims=1024)),
]


@pytest.mark.parametrize("name,kw", MODEL_FIXTURES)
def test_state_dict_keys_and_shapes_match_reference(golden, name, kw):
    g = golden(name)
    ref = dict(zip(g["keys"].tolist(), g["shapes"].tolist()))
    mine = {k: str(tuple(v.shape)) for k, v in PNV.PointNetVlad(**kw).state_dict().items()}
    assert sorted(mine) == sorted(ref)
    assert mine == ref
    # a reference-format checkpoint ({'epoch','iter','state_dict','optimizer','recall'}, train_pointnetvlad.py:180-186) loads strictly
    model = PNV.PointNetVlad(**kw)
    ckpt = {"epoch": 3, "iter": 7, "state_dict": synth.fill_state_dict({k: eval(s) for k, s in ref.items()}), "recall": np.float64(1.0)}
    model.load_state_dict(ckpt["state_dict"], strict=True)


def test_signatures_match_reference():
    def params(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values() if p.name != "self"]
    E = inspect.Parameter.empty
    assert params(PNV.PointNetVlad.__init__) == [("num_points", 4096), ("global_feat", True), ("feature_transform", False),
                                                 ("max_pool", False), ("output_dim", 256), ("emb_dims", 1024),
                                                 ("featnet", "lpdnet"), ("xyz_trans", False)]
    assert params(PNV.NetVLADLoupe.__init__) == [("feature_size", E), ("max_samples", E), ("cluster_size", E), ("output_dim", E),
                                                 ("gating", True), ("add_batch_norm", True), ("is_training", True)]
    assert params(PNV.GatingContext.__init__) == [("dim", E), ("add_batch_norm", True)]
    assert params(PNV.STN3d.__init__) == [("num_points", 2500), ("k", 3), ("use_bn", True)]
    assert params(PNV.PointNetfeat.__init__) == [("num_points", 2500), ("global_feat", True), ("feature_transform", False),
                                                 ("max_pool", True), ("emb_dims", 1024)]
    for cls in (LM.LPDNet, LM.LPDNetOrign):
        assert params(cls.__init__) == [("emb_dims", 512), ("use_mFea", False), ("t3d", True), ("tfea", False), ("use_relu", False)]
    assert params(LM.TranformNet.__init__) == [("k", 3), ("negative_slope", 1e-2), ("use_relu", True)]
    assert params(LM.knn) == [("x", E), ("k", E)]
    assert params(LM.get_graph_feature) == [("x", E), ("k", 20), ("idx", None)]
    assert params(LM.get_graph_feature_Origin) == [("x", E), ("k", 20), ("idx", None), ("cat", True)]
    assert params(L.quadruplet_loss) == [("q_vec", E), ("pos_vecs", E), ("neg_vecs", E), ("other_neg", E), ("m1", E), ("m2", E),
                                         ("use_min", False), ("lazy", False), ("ignore_zero_loss", False)]
    assert params(L.triplet_loss) == [("q_vec", E), ("pos_vecs", E), ("neg_vecs", E), ("margin", E),
                                      ("use_min", False), ("lazy", False), ("ignore_zero_loss", False)]
    assert params(L.best_pos_distance) == [("query", E), ("pos_vecs", E)]
    assert LM.LPDNet().k == 20 and LM.LPDNetOrign().k == 20  # mutable attribute, not a ctor argument


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "lpd_b200.h").read_text()
    declared = set(re.findall(r"\b(lpd_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lpd_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes binding and header disagree"
    assert lib.lpd_abi_version() == _lib.ABI_VERSION
    assert lib.lpd_status_str(-1).decode() == "invalid argument"


def test_no_cpu_fallback():
    model = PNV.PointNetVlad(num_points=64, featnet="lpdnet").eval()
    with pytest.raises(_lib.LpdError):
        model(torch.rand(2, 1, 64, 3))
    with pytest.raises(_lib.LpdError):
        LM.knn(torch.rand(1, 3, 64), 4)
    with pytest.raises(_lib.LpdError):
        L.quadruplet_loss(torch.rand(2, 1, 8), torch.rand(2, 2, 8), torch.rand(2, 3, 8), torch.rand(2, 1, 8), 0.5, 0.2)
    with pytest.raises(ValueError):
        PNV.PointNetVlad(featnet="nope")


def test_missing_library_fails_loudly():
    code = ("import os, sys; os.environ['LPD_B200_LIB']='/nonexistent/liblpd.so'; sys.path.insert(0, %r);"
            "from lpdnet_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept _lib.LpdError as e:\n    print('LOUD', e)\n" % str(ROOT))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "LOUD" in out.stdout and "no CPU or torch fallback" in out.stdout


def test_synthetic_database_structure():
    DB, Q, SETS = synth.descriptor_database(runs=3)
    assert len(DB) == 3 and DB[0].shape == (956, 256) and Q[1].shape == (132, 256)
    assert any(len(SETS[0][i][1]) == 0 for i in range(132)) and any(len(SETS[0][i][1]) == 3 for i in range(132))
