"""SURVEY §8f "next" rows: submap file format, hard-negative mining, feature representation / update_vectors."""
import numpy as np
import pytest
import torch

from lpdnet_b200 import loading_pointclouds as lp


def test_bin_loader_round_trip_and_wrong_size(tmp_path):
    """reference loading_pointclouds.py:26-47: float64 raw file of 4096 x 3; other sizes are reported and skipped"""
    r = np.random.default_rng(5)
    good = r.uniform(-1, 1, (4096, 3))
    bad = r.uniform(-1, 1, (1000, 3))
    good.astype(np.float64).tofile(tmp_path / "a.bin")
    bad.astype(np.float64).tofile(tmp_path / "b.bin")
    (good * 2).astype(np.float64).tofile(tmp_path / "c.bin")
    pc = lp.load_pc_file("a.bin", str(tmp_path))
    assert pc.dtype == np.float64 and pc.shape == (4096, 3) and np.array_equal(pc, good)
    assert lp.load_pc_file("b.bin", str(tmp_path)).shape == (0,)
    pcs = lp.load_pc_files(["a.bin", "b.bin", "c.bin"], str(tmp_path))
    assert pcs.shape == (2, 4096, 3) and np.array_equal(pcs[1], good * 2)
    x = lp.to_model_input(pcs)
    assert x.dtype == np.float32 and x.shape == (2, 1, 4096, 3) and x.flags["C_CONTIGUOUS"]
    assert np.array_equal(x[0, 0], good.astype(np.float32))


@pytest.mark.gpu
def test_hard_negative_mining_matches_kdtree(cuda):
    """reference util/data.py:103-115: KDTree over the cached descriptors of the sampled negatives"""
    from sklearn.neighbors import KDTree
    from lpdnet_b200.util import data
    r = np.random.default_rng(11)
    table = r.standard_normal((5000, 256)).astype(np.float32)
    table /= np.linalg.norm(table, axis=1, keepdims=True)
    for trial in range(4):
        negs = r.choice(5000, size=2000, replace=False).tolist()
        q = table[r.integers(0, 5000)] + 0.05 * r.standard_normal(256).astype(np.float32)
        want = np.squeeze(np.array(negs)[KDTree(table[negs]).query(np.array([q]), k=10)[1][0]]).tolist()
        assert data.get_random_hard_negatives(q, negs, 10, latent_vectors=table) == want
        data.TRAINING_LATENT_VECTORS = table          # the reference's global
        assert data.get_random_hard_negatives(q, negs, 10) == want
    qs = table[:3] + 0.05 * r.standard_normal((3, 256)).astype(np.float32)
    lists = [r.choice(5000, size=500, replace=False).tolist() for _ in range(3)]
    got = data.hard_negatives_batch(qs, torch.from_numpy(table), lists, 7)
    for i in range(3):
        want = np.array(lists[i])[KDTree(table[lists[i]]).query(qs[i:i + 1], k=7)[1][0]].tolist()
        assert got[i] == want


@pytest.mark.gpu
def test_feature_representation_and_update_vectors(cuda):
    """reference util/data.py:117-133, :277-354: single-submap and bulk embedding agree with the batched model call and
    leave the model in train mode"""
    from lpdnet_b200 import synth
    from lpdnet_b200.util import data
    from lpdnet_b200.util.PointNetVlad import PointNetVlad
    model = PointNetVlad(num_points=1024, featnet="lpdnet", emb_dims=1024)
    model.load_state_dict(synth.synthetic_state_dict(model))
    model = model.cuda().eval()
    clouds = synth.clouds(5, 1024)[:, 0].numpy()
    with torch.no_grad():
        want = model(torch.from_numpy(clouds).unsqueeze(1).cuda()).cpu().numpy()
    one = data.get_feature_representation(clouds[2], model)
    assert model.training and one.shape == (256,)
    assert np.abs(one - want[2]).max() < 1e-5
    model.eval()
    vecs = data.update_vectors(model, clouds, batch_num=2)      # ragged tail batch
    assert model.training and vecs.shape == (5, 256) and data.TRAINING_LATENT_VECTORS is vecs
    assert np.abs(vecs - want).max() < 1e-5


def test_checkpoint_format_and_scheduler(tmp_path):
    """reference train_pointnetvlad.py:64-76,92,172-199: checkpoint dict keys, best-copy rule, both load paths, LR policy
    (host logic; a CPU model is enough: the modules construct and (de)serialise without a device)"""
    from lpdnet_b200 import train_pointnetvlad as tp
    from lpdnet_b200.util.PointNetVlad import PointNetVlad
    model = PointNetVlad(num_points=256, featnet="lpdnetorigin", emb_dims=128)
    opt = torch.optim.Adam(model.parameters(), 1e-3)
    best = tp.save_model(model, opt, 3, 1234, 71.5, str(tmp_path))
    assert best == 71.5 and (tmp_path / "3-model.ckpt").exists() and (tmp_path / "best-model.ckpt").exists()
    best = tp.save_model(model, opt, 4, 2000, 60.0, str(tmp_path), best_so_far=best)
    ck = torch.load(tmp_path / "best-model.ckpt", weights_only=False)
    assert best == 71.5 and ck["epoch"] == 3                                   # a worse epoch does not replace the best copy
    assert set(ck) == {"epoch", "iter", "state_dict", "optimizer", "recall"}
    other = PointNetVlad(num_points=256, featnet="lpdnetorigin", emb_dims=128)
    opt2 = torch.optim.Adam(other.parameters(), 5e-4)
    assert tp.load_checkpoint(other, opt2, str(tmp_path / "4-model.ckpt")) == (5, 2000)
    for (ka, va), (kb, vb) in zip(model.state_dict().items(), other.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    assert opt2.param_groups[0]["lr"] == 1e-3
    torch.save(model.state_dict(), tmp_path / "weights.t7")                    # bare state_dict path (strict=False)
    third = PointNetVlad(num_points=256, featnet="lpdnetorigin", emb_dims=128)
    assert tp.load_checkpoint(third, None, str(tmp_path / "weights.t7")) == (0, 0)
    assert torch.equal(third.net_vlad.cluster_weights, model.net_vlad.cluster_weights)
    assert tp.load_checkpoint(third, None, str(tmp_path / "missing.ckpt")) == (0, 0)
    sched = tp.make_scheduler(opt)
    for recall in (50.0, 50.0, 50.0, 50.0):                                    # no improvement for > patience epochs
        sched.step(recall)
    assert abs(opt.param_groups[0]["lr"] - 2e-4) < 1e-12


def test_epoch_loop_schedule_on_cpu(tmp_path, monkeypatch):
    """reference train_pointnetvlad.py:38-170: the loop's schedule — base loader up to epoch 7, the
    hard-negative loader afterwards with a descriptor refresh at the switch and every 700 (epoch + 1) samples, evaluation +
    checkpoint + ReduceLROnPlateau after every epoch, resume from a checkpoint.  Host logic only: the step itself is stubbed."""
    from lpdnet_b200 import train_pointnetvlad as tp
    calls = {"base": 0, "advance": 0, "update": 0, "eval": 0, "iters": []}
    model = torch.nn.Linear(4, 4)

    def fake_step(model_, optimizer, q, p, n, o, m1, m2, **kw):
        assert (m1, m2) == (0.5, 0.2) and kw["use_min"] and kw["lazy"] and not kw["ignore_zero_loss"]
        calls[q] += 1
        return torch.tensor(0.25)

    monkeypatch.setattr(tp, "train_step", fake_step)
    base = [("base",) * 4] * 3
    advance = [("advance",) * 4] * 800                       # 1600 samples per epoch: crosses 700 * (epoch + 1) multiples

    def update():
        calls["update"] += 1

    def evaluate_fn(m):
        calls["eval"] += 1
        return 10.0, 0.5, 40.0 + calls["eval"]

    logged = []
    cfg = tp.TrainConfig(max_epoch=10, optimizer="momentum", model_save_path=str(tmp_path))
    state = tp.train(model, base, advance, evaluate_fn, cfg, update_vectors=update, log=lambda n, v, i: logged.append((n, i)))
    assert calls["base"] == 3 * 8 and calls["advance"] == 800 * 2 and calls["eval"] == 10
    # refreshes: one at the switch (epoch 8), then whenever TOTAL_ITERATIONS hits a multiple of 700 * (epoch + 1) (rounded to the batch)
    iters, expect = 3 * 8 * 2, 1
    for epoch in (8, 9):
        for _ in range(800):
            iters += 2
            expect += iters % (700 * (epoch + 1) // 2 * 2) == 0
    assert calls["update"] == expect and state["iter"] == iters and state["epoch"] == 9
    assert (tmp_path / "9-model.ckpt").exists() and (tmp_path / "best-model.ckpt").exists() and state["best"] == 50.0
    assert ("Val Recall", 9) in logged and ("Loss", 0) in logged
    # resume: starts at the epoch after the checkpoint's, keeps the iteration counter
    calls.update(base=0, advance=0, update=0, eval=0)
    cfg2 = tp.TrainConfig(max_epoch=11, optimizer="momentum", model_save_path=str(tmp_path), pretrained_path=str(tmp_path / "9-model.ckpt"))
    state2 = tp.train(model, base, advance, evaluate_fn, cfg2, update_vectors=update)
    assert calls["base"] == 0 and calls["advance"] == 800 and state2["epoch"] == 10 and state2["iter"] == iters + 1600
    assert calls["update"] >= 1                               # starting_epoch > DIVISION_EPOCH + 1: refresh before the first step
