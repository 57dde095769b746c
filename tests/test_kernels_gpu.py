"""Kernel-level parity: every C-ABI entry point against the CPU oracle / a float64 numpy restatement,
on seeded inputs, including ragged and edge shapes.  Runs on the B200 box (-m gpu)."""
import numpy as np
import pytest
import torch

from lpdnet_b200 import ops, synth
from oracle import knn_canonical, loss_numpy, retrieval_bruteforce

pytestmark = pytest.mark.gpu


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def rng(seed):
    return np.random.default_rng(seed)


# ------------------------------------------------------------------------------------------------ kNN
@pytest.mark.parametrize("B,N,C,k", [
    (2, 1024, 3, 20), (2, 512, 64, 20), (1, 1000, 3, 20), (3, 130, 64, 20), (1, 257, 3, 32),
    (1, 64, 8, 1), (1, 33, 16, 32), (2, 200, 5, 7), (1, 4096, 3, 20), (1, 2048, 64, 32),
])
def test_knn_bit_exact_vs_canonical_oracle(cuda, B, N, C, k):
    r = rng(B * 1000 + N + C)
    x = (r.uniform(-1, 1, (B, N, C)) if C <= 8 else np.maximum(r.standard_normal((B, N, C)), 0.01 * r.standard_normal((B, N, C)))).astype(np.float32)
    want = knn_canonical(x, k)
    got = ops.knn(dev(x), k).cpu().numpy()
    assert got.dtype == np.int32
    assert np.array_equal(got, want)
    got64 = ops.knn(dev(x), k, int64=True).cpu().numpy()
    assert got64.dtype == np.int64 and np.array_equal(got64, want)


def test_knn_exact_ties_and_duplicates(cuda):
    lat = np.stack(np.meshgrid(np.arange(8.), np.arange(8.), np.arange(8.), indexing="ij"), 0).reshape(3, 512).T
    x = synth.clouds(1, 1024)[0, 0].numpy().copy()
    x[500:510] = x[0:10]  # exact duplicate points
    for cloud in (lat.astype(np.float32), x):
        want = knn_canonical(cloud[None], 20)
        got = ops.knn(dev(cloud[None]), 20).cpu().numpy()
        assert np.array_equal(got, want)


@pytest.mark.parametrize("kind", ["clustered", "planar", "offset", "line", "identical", "lattice16", "outliers", "tiny"])
@pytest.mark.parametrize("k", [20, 7, 32])
def test_knn_xyz_grid_is_bit_exact_on_adversarial_clouds(cuda, kind, k):
    """lpd_knn_xyz (grid-pruned search) must return exactly the canonical lists whatever the point distribution: the grid
    only prunes, with a bound that covers the fp32 rounding of the canonical score."""
    r = rng(len(kind) * 7 + k)
    N = 2048
    if kind == "clustered":        # LiDAR-like: dense blobs + sparse background
        x = np.concatenate([r.normal(c, 0.01, (400, 3)) for c in r.uniform(-1, 1, (4, 3))] + [r.uniform(-1, 1, (N - 1600, 3))])
    elif kind == "planar":         # ground plane: z extent ~ 0
        x = np.concatenate([r.uniform(-1, 1, (N, 2)), r.normal(0, 1e-4, (N, 1))], 1)
    elif kind == "offset":         # far from the origin: the expansion -xx + 2dot - xx cancels heavily
        x = r.uniform(-1, 1, (N, 3)) + np.array([100.0, -50.0, 25.0])
    elif kind == "line":           # degenerate extent on two axes
        x = np.concatenate([r.uniform(-1, 1, (N, 1)), np.zeros((N, 2))], 1)
    elif kind == "identical":      # every point the same: all ties, order = index
        x = np.tile(r.uniform(-1, 1, (1, 3)), (N, 1))
    elif kind == "lattice16":      # masses of exact ties across cell boundaries
        g = np.arange(16.0) / 8 - 1
        x = np.stack(np.meshgrid(g, g, g[:8], indexing="ij"), -1).reshape(-1, 3)
    elif kind == "outliers":       # a few far points stretch the bounding box: almost everything lands in one cell
        x = np.concatenate([r.uniform(-0.01, 0.01, (N - 4, 3)), r.uniform(-50, 50, (4, 3))])
    else:                          # N below the grid path threshold and below L
        x = r.uniform(-1, 1, (70, 3))
    x = x.astype(np.float32)[None]
    want = knn_canonical(x, k)
    got = ops.knn(dev(x), k).cpu().numpy()
    assert np.array_equal(got, want)
    prev, ops.KNN_GRID = ops.KNN_GRID, False
    try:
        assert np.array_equal(ops.knn(dev(x), k).cpu().numpy(), want)       # brute-force kernel: same lists
    finally:
        ops.KNN_GRID = prev


def test_knn_xyz_on_the_grid_left_by_cell_order(cuda):
    """knn(xyz_sorted) right after cell_order() reuses that call's grid (lpd_cell_order_grid + lpd_knn_xyz_ordered): same lists as the
    brute-force kernel on the re-ordered cloud; an edited or foreign cloud takes the ordinary path"""
    for B, N, k in ((3, 1000, 20), (2, 4096, 20), (2, 700, 9), (1, 2048, 32)):
        x = synth.clouds(B, N, seed=5 + N)[:, 0].contiguous().cuda()
        perm, inv, xs = ops.cell_order(x, want_inv=True)
        assert ops._grid_of(xs) is not None
        got = ops.knn(xs, k).cpu().numpy()
        assert np.array_equal(got, knn_canonical(xs.cpu().numpy(), k))
        assert np.array_equal(xs.cpu().numpy(), np.take_along_axis(x.cpu().numpy(), perm.cpu().numpy()[..., None].astype(np.int64), 1))
        other = xs.clone()
        assert ops._grid_of(other) is None
        assert np.array_equal(ops.knn(other, k).cpu().numpy(), got)
        xs.mul_(1.5)                                                        # edited in place: the cached grid no longer applies
        assert ops._grid_of(xs) is None
        assert np.array_equal(ops.knn(xs, k).cpu().numpy(), knn_canonical(xs.cpu().numpy(), k))


def test_knn_matches_reference_golden_sets(cuda, golden):
    """against the reference's own torch knn() output (tie-aware: sets may differ only on near-ties)"""
    from _helpers import assert_knn_equivalent
    g = golden("knn")
    x = synth.clouds(2, 1024)[:, 0].contiguous()
    got = ops.knn(x.cuda(), 20).cpu().numpy()
    assert assert_knn_equivalent(x.numpy(), got, g["xyz_idx"].astype(np.int64), 20) <= 8


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(130, 70, 19), (64, 64, 64), (257, 128, 3), (1000, 1024, 512), (5, 256, 1024), (300, 64, 1024)])
@pytest.mark.parametrize("al", [ops.A_MK, ops.A_KM])
@pytest.mark.parametrize("bl", [ops.B_NK, ops.B_KN])
def test_gemm_layouts_and_ragged_shapes(cuda, M, N, K, al, bl):
    r = rng(M + 7 * N + 13 * K)
    A = r.standard_normal((M, K)).astype(np.float32)
    Bm = r.standard_normal((K, N)).astype(np.float32)
    scale = r.standard_normal(N).astype(np.float32)
    shift = r.standard_normal(N).astype(np.float32)
    ref = (A.astype(np.float64) @ Bm.astype(np.float64)) * scale + shift
    ref = np.where(ref > 0, ref, 0.01 * ref)
    a_dev = dev(A if al == ops.A_MK else A.T)
    b_dev = dev(Bm.T if bl == ops.B_NK else Bm)
    out = ops.gemm(a_dev, b_dev, a_layout=al, b_layout=bl, M=M, N=N, K=K, scale=dev(scale), shift=dev(shift),
                   act=ops.ACT_LEAKY, slope=0.01).cpu().numpy()
    assert np.abs(out - ref).max() < max(2e-6 * np.abs(ref).max(), 1e-5)


def test_gemm_batched_strided_gate_and_ldc(cuda):
    r = rng(5)
    b, M, N, K = 3, 70, 96, 40
    A = r.standard_normal((b, M, K)).astype(np.float32)
    Bm = r.standard_normal((b, N, K)).astype(np.float32)
    aux = r.standard_normal((b, M, N)).astype(np.float32)
    ref = np.einsum("bmk,bnk->bmn", A.astype(np.float64), Bm.astype(np.float64))
    ref = aux / (1 + np.exp(-ref))
    out = ops.gemm(dev(A), dev(Bm), M=M, N=N, K=K, batch=b, strideA=M * K, strideB=N * K, act=ops.ACT_GATE, aux=dev(aux))
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-4
    # write into a column window of a wider buffer (the concat-free layout of the feature pyramid)
    wide = torch.zeros(M, 256, device="cuda")
    ops.gemm(dev(A[0]), dev(Bm[0]), M=M, N=N, K=K, out=wide[:, 128:], ldc=256)
    ref0 = A[0].astype(np.float64) @ Bm[0].astype(np.float64).T
    assert np.abs(wide[:, 128:128 + N].cpu().numpy() - ref0).max() < 1e-4
    assert float(wide[:, :128].abs().max()) == 0.0 and float(wide[:, 128 + N:].abs().max()) == 0.0


def test_bn_fold_transpose_colmax_splitk(cuda):
    r = rng(11)
    C = 130
    g, b, m = (r.standard_normal(C).astype(np.float32) for _ in range(3))
    v = r.uniform(0.5, 2, C).astype(np.float32)
    bias = r.standard_normal(C).astype(np.float32)
    s, t = ops.bn_fold(dev(g), dev(b), dev(m), dev(v), 1e-5, bias=dev(bias))
    s_ref = g / np.sqrt(v.astype(np.float64) + 1e-5)
    assert np.abs(s.cpu().numpy() - s_ref).max() < 1e-6
    assert np.abs(t.cpu().numpy() - (b - m * s_ref + bias * s_ref)).max() < 1e-5
    x = r.standard_normal((3, 77, 45)).astype(np.float32)
    assert np.array_equal(ops.transpose(dev(x)).cpu().numpy(), x.transpose(0, 2, 1))
    assert np.array_equal(ops.colmax(dev(x), 3, 77, 45).cpu().numpy(), x.max(1))
    part = r.standard_normal((9, 6, 20)).astype(np.float32)
    out = ops.splitk_reduce(dev(part), 9, 6, 20, scale=dev(g[:20]), shift=dev(b[:20])).cpu().numpy()
    assert np.abs(out - (part.astype(np.float64).sum(0) * g[:20] + b[:20])).max() < 1e-5


@pytest.mark.parametrize("M,N,K,lda,out_half", [(300, 64, 1024, 1024, False), (1000, 1024, 512, 512, True), (129, 512, 128, 256, False),
                                                (4096, 256, 64, 64, True), (128, 68, 72, 80, False),
                                                (70000, 512, 128, 128, True), (257, 256, 200, 208, True), (40000, 1024, 512, 512, True)])
def test_gemm_f16_operands(cuda, M, N, K, lda, out_half):
    """lpd_gemm_f16: fp16 operands, fp32 accumulation on tcgen05 kind::f16, fused affine + LeakyReLU epilogue, fp32 or fp16
    output — against float64 on the SAME fp16-rounded operands (so only accumulation order and the output rounding differ)"""
    r = rng(M + N + K)
    A = (r.standard_normal((M, lda))).astype(np.float16)
    W = (r.standard_normal((N, K)) / np.sqrt(K)).astype(np.float16)
    sc, sh = r.standard_normal(N).astype(np.float32), r.standard_normal(N).astype(np.float32)
    ref = A[:, :K].astype(np.float64) @ W.astype(np.float64).T * sc + sh
    ref = np.where(ref > 0, ref, 0.01 * ref)
    dA, dW = torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda()
    got = ops.gemm_f16(dA, dW, M=M, N=N, K=K, lda=lda, out_half=out_half, scale=dev(sc), shift=dev(sh), act=ops.ACT_LEAKY, slope=0.01)
    assert got.dtype == (torch.float16 if out_half else torch.float32)
    tol = (2.0 ** -10 if out_half else 1e-5) * max(1.0, np.abs(ref).max())
    assert np.abs(got.float().cpu().numpy() - ref).max() <= tol
    # round trip of the conversion kernel
    x32 = r.standard_normal((37, 50)).astype(np.float32)
    assert np.array_equal(ops.to_f16(dev(x32)).cpu().numpy(), x32.astype(np.float16))


@pytest.mark.parametrize("M,N,K", [(1000, 256, 64), (70000, 256, 64), (300, 512, 100), (513, 128, 64)])
def test_gemm_tf32_with_fp16_output(cuda, M, N, K):
    """lpd_gemm_tf32_out16 (fp32 operands as TF32, fp16 output; N % 256 == 0 runs on the CTA-pair kernel) against float64 on the
    TF32-truncated operands"""
    r = rng(M + N + K)
    A = r.standard_normal((M, K)).astype(np.float32)
    W = (r.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    sc, sh = r.standard_normal(N).astype(np.float32), r.standard_normal(N).astype(np.float32)
    trunc = lambda x: (x.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
    ref = trunc(A).astype(np.float64) @ trunc(W).astype(np.float64).T * sc + sh
    ref = np.maximum(ref, 0.0)
    got = ops.gemm_tf32_out16(dev(A), dev(W), M=M, N=N, K=K, scale=dev(sc), shift=dev(sh), act=ops.ACT_RELU)
    assert got.dtype == torch.float16 and tuple(got.shape) == (M, N)
    assert np.abs(got.float().cpu().numpy() - ref).max() <= 2.0 ** -10 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("M,N,K,batch", [(1024, 64, 4096, 3), (128, 64, 64, 1), (200, 72, 1000, 1), (1024, 64, 16384, 2)])
def test_gemm_f16_tn_contraction_over_rows(cuda, M, N, K, batch):
    """lpd_gemm_f16_tn: out[z] = A[z]^T . B[z] over the rows of two fp16 point-major maps (the NetVLAD aggregate in f16 mode)"""
    r = rng(M + N + K + batch)
    A = r.standard_normal((batch * K, M)).astype(np.float16)
    Bm = r.random((batch * K, N)).astype(np.float16)
    got = ops.gemm_f16_tn(torch.from_numpy(A).cuda(), torch.from_numpy(Bm).cuda(), M=M, N=N, K=K, lda=M, ldb=N, batch=batch).cpu().numpy()
    for z in range(batch):
        ref = A[z * K:(z + 1) * K].astype(np.float64).T @ Bm[z * K:(z + 1) * K].astype(np.float64)
        assert np.abs(got[z] - ref).max() <= 2e-6 * K ** 0.5 * max(1.0, np.abs(ref).max()), f"slice {z}"


# ------------------------------------------------------------------------------------------------ EdgeConv
def _leaky(x):
    return np.where(x > 0, x, 0.01 * x)


@pytest.mark.parametrize("B,N,k,C", [(2, 300, 20, 256), (1, 257, 7, 128), (2, 128, 32, 64)])
def test_edge_gather_ext_vs_materialised_edges(cuda, B, N, k, C):
    r = rng(N + C)
    p = r.standard_normal((B * N, C)).astype(np.float32)
    q = r.standard_normal((B * N, C)).astype(np.float32)
    idx = r.integers(0, N, (B, N, k)).astype(np.int32)
    s = r.standard_normal(C).astype(np.float32)  # negative scales included (min branch)
    t = r.standard_normal(C).astype(np.float32)
    pb = p.reshape(B, N, C)
    edges = pb[np.arange(B)[:, None, None], idx] + q.reshape(B, N, 1, C)      # [B,N,k,C]
    ref = _leaky(edges.astype(np.float64) * s + t).max(2).reshape(B * N, C)
    out = torch.empty(B * N, C, device="cuda")
    ops.edge_gather_ext(dev(p), C, dev(q), C, dev(idx, torch.int32), B, N, k, C, dev(s), dev(t), ops.ACT_LEAKY, 0.01, out, C)
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-5
    # gather-only edges (LPDNetOrign, cat=False)
    ref2 = pb[np.arange(B)[:, None, None], idx].max(2).reshape(B * N, C)
    ops.edge_gather_ext(dev(p), C, None, 0, dev(idx, torch.int32), B, N, k, C, None, None, ops.ACT_NONE, 0.0, out, C)
    assert np.array_equal(out.cpu().numpy(), ref2)


@pytest.mark.parametrize("B,N,k,C", [(2, 300, 20, 256), (1, 257, 7, 128), (3, 1000, 32, 256), (1, 129, 1, 128), (2, 64, 21, 256)])
def test_edge_gather_prescaled_wide_rows(cuda, B, N, k, C):
    """the pre-scaled form (scale = shift = None, C in {128, 256}: one warp per point, max only): out = act(q + max_m p_j),
    bit-exact against numpy (an fp32 max and one fp32 add), with a leading dimension wider than C like the model's buffers"""
    r = rng(N + C + k)
    ld = 2 * C
    pq = r.standard_normal((B * N, ld)).astype(np.float32)
    idx = r.integers(0, N, (B, N, k)).astype(np.int32)
    P, Q = pq[:, :C].reshape(B, N, C), pq[:, C:]
    want = _leaky_f32(P[np.arange(B)[:, None, None], idx].max(2).reshape(B * N, C) + Q)
    d_pq = dev(pq)
    out = torch.full((B * N, ld), -7.0, device="cuda")
    ops.edge_gather_ext(d_pq, ld, d_pq[:, C:], ld, dev(idx, torch.int32), B, N, k, C, None, None, ops.ACT_LEAKY, 0.01, out[:, C:], ld)
    got = out.cpu().numpy()
    assert np.array_equal(got[:, C:], want) and (got[:, :C] == -7.0).all()


def _leaky_f32(x):
    x = x.astype(np.float32)
    return np.where(x > 0, x, x * np.float32(0.01)).astype(np.float32)


@pytest.mark.parametrize("B,N,k,C", [(2, 300, 20, 128), (1, 257, 32, 128), (1, 100, 7, 128), (2, 200, 20, 64), (1, 90, 25, 64)])
def test_edgeconv_dg_vs_materialised_edges(cuda, B, N, k, C):
    r = rng(N + k + C)
    pq = r.standard_normal((B * N, 2 * C)).astype(np.float32)
    idx = r.integers(0, N, (B, N, k)).astype(np.int32)
    s1, t1, s2, t2 = (r.standard_normal(C).astype(np.float32) for _ in range(4))
    w2 = (r.standard_normal((C, C)) / np.sqrt(C)).astype(np.float32)
    P, Q = pq[:, :C].reshape(B, N, C), pq[:, C:].reshape(B, N, C)
    y1 = _leaky((P[np.arange(B)[:, None, None], idx] + Q[:, :, None, :]).astype(np.float64) * s1 + t1)   # [B,N,k,C]
    y2 = _leaky((y1 @ w2.astype(np.float64).T) * s2 + t2)
    x1_ref, x2_ref = y1.max(2).reshape(B * N, C), y2.max(2).reshape(B * N, C)
    d_pq = dev(pq)
    x = torch.zeros(B * N, 2 * C, device="cuda")
    ops.edgeconv_dg(d_pq, 2 * C, d_pq[:, C:], 2 * C, dev(idx, torch.int32), B, N, k, C, C, dev(s1), dev(t1), dev(w2), dev(s2), dev(t2),
                    ops.ACT_LEAKY, 0.01, x, 2 * C, x[:, C:], 2 * C)
    got = x.cpu().numpy()
    assert np.abs(got[:, :C] - x1_ref).max() < 1e-5
    assert np.abs(got[:, C:] - x2_ref).max() < 5e-5


# ------------------------------------------------------------------------------------------------ NetVLAD
def test_netvlad_assign_and_finish(cuda):
    r = rng(3)
    B, N, D, K = 2, 300, 256, 64
    x = r.standard_normal((B * N, D)).astype(np.float32)
    wc = (r.standard_normal((D, K)) / np.sqrt(D)).astype(np.float32)
    wc2 = (r.standard_normal((D, K)) / np.sqrt(D)).astype(np.float32)
    s, t = r.standard_normal(K).astype(np.float32), r.standard_normal(K).astype(np.float32)
    z = (x.astype(np.float64) @ wc) * s + t
    a_ref = np.exp(z - z.max(1, keepdims=True))
    a_ref /= a_ref.sum(1, keepdims=True)
    a = ops.netvlad_assign(dev(x), B * N, D, dev(wc), dev(s), dev(t))
    assert np.abs(a.cpu().numpy() - a_ref).max() < 1e-5
    # aggregate with lpd_gemm (A = x^T stored [K=N][M=D], B = a stored [K=N][N=K]) then finish
    vraw = ops.gemm(dev(x), a, a_layout=ops.A_KM, b_layout=ops.B_KN, M=D, N=K, K=N, lda=D, ldb=K, batch=B,
                    strideA=N * D, strideB=N * K)
    a3, x3 = a_ref.reshape(B, N, K), x.reshape(B, N, D).astype(np.float64)
    v = np.einsum("bnd,bnk->bdk", x3, a3) - a3.sum(1)[:, None, :] * wc2[None]
    assert np.abs(vraw.cpu().numpy() - np.einsum("bnd,bnk->bdk", x3, a3)).max() < 1e-4
    v = v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)
    v = v.reshape(B, D * K)
    v = v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)
    out = ops.netvlad_finish(vraw, a, dev(wc2), B, N, D, K)
    assert np.abs(out.cpu().numpy() - v).max() < 1e-6


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("B,N,D", [(3, 256, 1024), (2, 96, 128), (5, 160, 512)])
def test_netvlad_fused_assign_softmax_and_cluster_finish(cuda, f16, B, N, D):
    """lpd_gemm_softmax64 (softmax + 32-row partial sums as the GEMM epilogue, ragged last tile) and lpd_netvlad_finish_parts (cluster
    kernel) against a float64 restatement of PointNetVlad.py:48-74"""
    r = rng(11)
    K, M = 64, B * N
    x = r.standard_normal((M, D)).astype(np.float32)
    wc = (r.standard_normal((D, K)) / np.sqrt(D)).astype(np.float32)
    wc2 = (r.standard_normal((D, K)) / np.sqrt(D)).astype(np.float32)
    s, t = (1.0 + 0.3 * r.standard_normal(K)).astype(np.float32), r.standard_normal(K).astype(np.float32)
    if f16:
        xd, wd = dev(x).half(), dev(wc.T.copy()).half()
        xr, wr = xd.float().cpu().numpy().astype(np.float64), wd.float().cpu().numpy().astype(np.float64).T
        tol = 2e-6
    else:
        xd, wd = dev(x), dev(wc.T.copy())
        xr, wr = x.astype(np.float64), wc.astype(np.float64)
        tol = 2e-3          # TF32 operands
    z = (xr @ wr) * s + t
    a_ref = np.exp(z - z.max(1, keepdims=True))
    a_ref /= a_ref.sum(1, keepdims=True)
    a32, a16, apart = ops.gemm_softmax64(xd, wd, M=M, K=D, scale=dev(s), shift=dev(t), want32=True, want16=True, want_parts=True)
    got = a32.cpu().numpy()
    assert np.abs(got - a_ref).max() < tol
    assert np.abs(got.sum(1) - 1.0).max() < 1e-5
    assert np.abs(a16.float().cpu().numpy() - got).max() <= 2.0 ** -11          # fp16 copy of the same values
    nb = (M + 31) // 32
    want_parts = np.stack([got[i * 32:(i + 1) * 32].astype(np.float64).sum(0) for i in range(nb)])
    assert np.abs(apart.cpu().numpy() - want_parts).max() < 1e-5
    # finish on the partials == finish on the assignment (old path) == float64 restatement
    a3 = got.reshape(B, N, K).astype(np.float64)
    vraw = np.einsum("bnd,bnk->bdk", x.reshape(B, N, D).astype(np.float64), a3)
    v = vraw - a3.sum(1)[:, None, :] * wc2[None]
    v = v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)
    v = v.reshape(B, D * K)
    v = v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)
    out = ops.netvlad_finish_parts(dev(vraw.astype(np.float32)), apart, N // 32, dev(wc2), B, D, K)
    assert np.abs(out.cpu().numpy() - v).max() < 1e-6
    out2 = ops.netvlad_finish(dev(vraw.astype(np.float32)), a32, dev(wc2), B, N, D, K)
    assert np.abs(out2.cpu().numpy() - v).max() < 1e-6


def test_hidden_gate_matches_split_reduce_plus_gating(cuda):
    """lpd_hidden_gate == lpd_splitk_reduce + the gating GEMM (PointNetVlad.py:76-81, 103-115) against float64"""
    r = rng(12)
    for B, O, splits in ((64, 256, 128), (3, 100, 7), (1, 1024, 2)):
        part = r.standard_normal((splits, B, O)).astype(np.float32)
        s2, t2, sg, tg = (r.standard_normal(O).astype(np.float32) for _ in range(4))
        wg = (r.standard_normal((O, O)) / np.sqrt(O)).astype(np.float32)
        h = part.astype(np.float64).sum(0) * s2 + t2
        want = h / (1.0 + np.exp(-((h @ wg.astype(np.float64)) * sg + tg)))
        got = ops.hidden_gate(dev(part), splits, B, O, dev(s2), dev(t2), dev(wg), dev(sg), dev(tg)).cpu().numpy()
        assert np.abs(got - want).max() < 2e-5 * max(1.0, np.abs(want).max())
        got2 = ops.hidden_gate(dev(part), splits, B, O, None, None, dev(wg), None, dev(tg)).cpu().numpy()
        h2 = part.astype(np.float64).sum(0)
        want2 = h2 / (1.0 + np.exp(-((h2 @ wg.astype(np.float64)) + tg)))
        assert np.abs(got2 - want2).max() < 2e-5 * max(1.0, np.abs(want2).max())


# ------------------------------------------------------------------------------------------------ loss
def test_loss_forward_backward_vs_reference_golden(cuda, golden):
    g = golden("loss")
    for tag in ("2_2_18", "3_1_2", "5_4_7"):
        q, pos, neg, other = (dev(g[f"{tag}.{n}"]) for n in ("q", "pos", "neg", "other"))
        for um in (0, 1):
            for lz in (0, 1):
                for ig in (0, 1):
                    ft = f"{tag}.{um}{lz}{ig}"
                    loss, grads = ops.quadruplet_loss(q, pos, neg, other, 0.5, 0.2, um, lz, ig, need_grad=True)
                    ref = float(g[ft + ".quad"])
                    assert abs(float(loss) - ref) <= 1e-5 * abs(ref), ft  # north_star: 1e-5 relative
                    for a, n in zip(grads, ("gq", "gpos", "gneg", "gother")):
                        r_ = g[f"{ft}.quad.{n}"]
                        assert np.abs(a.cpu().numpy() - r_).max() <= 1e-5 * max(1.0, np.abs(r_).max()), (ft, n)
                    loss, grads = ops.quadruplet_loss(q, pos, neg, None, 0.5, 0.2, um, lz, ig, need_grad=True)
                    ref = float(g[ft + ".trip"])
                    assert abs(float(loss) - ref) <= 1e-5 * abs(ref), ft
                    for a, n in zip(grads[:3], ("gq", "gpos", "gneg")):
                        r_ = g[f"{ft}.trip.{n}"]
                        assert np.abs(a.cpu().numpy() - r_).max() <= 1e-5 * max(1.0, np.abs(r_).max()), (ft, n)
    # zero-loss tuples + ignore_zero_loss, checked against the numpy oracle
    r = rng(4)
    q = r.standard_normal((4, 1, 32)).astype(np.float32) * 0.1
    pos = q + 0.01 * r.standard_normal((4, 2, 32)).astype(np.float32)
    neg = q + 0.05 * r.standard_normal((4, 6, 32)).astype(np.float32)
    neg[1] += 5.0  # tuple 1 has zero loss
    other = q + 0.05 * r.standard_normal((4, 1, 32)).astype(np.float32)
    for lz in (0, 1):
        ref = loss_numpy.quadruplet_loss(q, pos, neg, other, 0.5, 0.2, True, bool(lz), True)
        got = ops.quadruplet_loss(dev(q), dev(pos), dev(neg), dev(other), 0.5, 0.2, 1, lz, 1)
        assert abs(float(got) - ref) <= 1e-5 * abs(ref)


# ------------------------------------------------------------------------------------------------ retrieval
@pytest.mark.parametrize("Ndb,Nq,D,k", [(956, 132, 256, 25), (21988, 300, 256, 25), (70, 5, 256, 25), (20, 3, 64, 25), (1000, 65, 30, 10)])
def test_retrieval_topk_vs_bruteforce_oracle(cuda, Ndb, Nq, D, k):
    r = rng(Ndb + Nq)
    db = r.standard_normal((Ndb, D)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    q = (db[r.integers(0, Ndb, Nq)] + 0.3 * r.standard_normal((Nq, D)) / np.sqrt(D)).astype(np.float32)
    if Ndb > 100:
        db[50] = db[7]  # exact duplicate rows: tie broken towards the lower index
    want_idx, want_d = retrieval_bruteforce(db, q, k)
    idx, dist = ops.retrieval_topk(dev(db), dev(q), k)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    fin = np.isfinite(want_d)
    assert np.array_equal(dist.cpu().numpy()[fin], want_d[fin])
    idx2, _ = ops.retrieval_topk(dev(db), dev(q), k, idx_offset=1000, want_dist=False)
    assert np.array_equal(idx2.cpu().numpy()[want_idx >= 0], want_idx[want_idx >= 0] + 1000)


@pytest.mark.parametrize("M,D,ld,act", [(1000, 3, 3, ops.ACT_LEAKY), (4096, 8, 8, ops.ACT_RELU), (130, 3, 5, ops.ACT_NONE), (70000, 3, 3, ops.ACT_LEAKY)])
def test_pointwise_mlp2_matches_two_fp32_layers(cuda, M, D, ld, act):
    """lpd_pointwise_mlp2 (conv1 + conv2 of the feature nets in one pass) == the two strict-fp32 layers, and == fp64 numpy"""
    r = rng(M + D)
    x = r.standard_normal((M, ld)).astype(np.float32)
    w1 = (r.standard_normal((64, D)) / np.sqrt(D)).astype(np.float32)
    w2 = (r.standard_normal((64, 64)) / 8).astype(np.float32)
    s1, t1, s2, t2 = (r.standard_normal(64).astype(np.float32) for _ in range(4))
    slope = 0.01

    def a(v):
        return v if act == ops.ACT_NONE else np.where(v > 0, v, (slope if act == ops.ACT_LEAKY else 0.0) * v)
    ref = a((a((x[:, :D].astype(np.float64) @ w1.astype(np.float64).T) * s1 + t1) @ w2.astype(np.float64).T) * s2 + t2)
    dx = dev(x)
    got = ops.pointwise_mlp2(dx, D, M, dev(w1), dev(s1), dev(t1), dev(w2), dev(s2), dev(t2), act, slope).cpu().numpy()
    assert np.abs(got - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
    h = ops.gemm(dx, dev(w1), M=M, N=64, K=D, lda=ld, scale=dev(s1), shift=dev(t1), act=act, slope=slope)
    two = ops.gemm(h, dev(w2), M=M, N=64, K=64, scale=dev(s2), shift=dev(t2), act=act, slope=slope).cpu().numpy()
    assert np.abs(got - two).max() < 2e-5 * max(1.0, np.abs(ref).max())


# ------------------------------------------------------------------------------------------------ tensor-core GEMM
@pytest.mark.parametrize("M,N,K,ldx", [(1000, 1024, 512, 0), (128, 64, 1024, 0), (300, 512, 128, 0), (4096, 256, 64, 0),
                                       (777, 200, 96, 0), (2048, 128, 128, 512), (130, 1024, 40, 0), (40000, 1024, 512, 0), (5001, 512, 100, 0)])
def test_gemm_tf32_tensor_core_path(cuda, M, N, K, ldx):
    """tcgen05 kind::tf32: operands are rounded to TF32 (10-bit mantissa) -> tolerance 2^-9 of the result scale."""
    r = rng(M + N + K)
    lda = ldx if ldx else K
    A = r.standard_normal((M, lda)).astype(np.float32)
    W = (r.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    scale, shift = r.standard_normal(N).astype(np.float32), r.standard_normal(N).astype(np.float32)
    ref = (A[:, :K].astype(np.float64) @ W.astype(np.float64).T) * scale + shift
    ref = np.where(ref > 0, ref, 0.01 * ref)
    wide = torch.zeros(M, N + 64, device="cuda")
    ops.gemm_tf32(dev(A), dev(W), M=M, N=N, K=K, lda=lda, out=wide[:, 32:], ldc=N + 64, scale=dev(scale), shift=dev(shift),
                  act=ops.ACT_LEAKY, slope=0.01)
    got = wide.cpu().numpy()
    assert np.abs(got[:, 32:32 + N] - ref).max() < 2.0 ** -9 * np.abs(ref).max()
    assert not got[:, :32].any() and not got[:, 32 + N:].any()        # nothing written outside the column window
    # and it agrees with the strict fp32 path to TF32 accuracy
    strict = ops.gemm(dev(A), dev(W), M=M, N=N, K=K, lda=lda, scale=dev(scale), shift=dev(shift), act=ops.ACT_LEAKY, slope=0.01)
    assert np.abs(strict.cpu().numpy() - got[:, 32:32 + N]).max() < 2.0 ** -9 * np.abs(ref).max()


@pytest.mark.parametrize("M,N,K,batch,accumulate", [(4096, 64, 1024, 3, False), (1024, 1024, 64, 4, True), (300, 200, 96, 2, True),
                                                     (128, 128, 512, 1, True), (640, 72, 40, 5, False)])
def test_gemm_tf32_batched_and_accumulating(cuda, M, N, K, batch, accumulate):
    """lpd_gemm_tf32_ex: row-stacked slices, optional C += ... (NetVLAD backward products, accumulating input gradients);
    slices whose M / N are not tile multiples must not leak into their neighbours."""
    r = rng(M + N + K + batch)
    A = r.standard_normal((batch * M, K)).astype(np.float32)
    W = (r.standard_normal((batch * N, K)) / np.sqrt(K)).astype(np.float32)
    C0 = r.standard_normal((batch * M, N)).astype(np.float32)
    out = dev(C0.copy())
    ops.gemm_tf32(dev(A), dev(W), M=M, N=N, K=K, batch=batch, out=out, ldc=N, accumulate=accumulate)
    got = out.cpu().numpy()
    for z in range(batch):
        ref = A[z * M:(z + 1) * M].astype(np.float64) @ W[z * N:(z + 1) * N].astype(np.float64).T
        if accumulate:
            ref = ref + C0[z * M:(z + 1) * M]
        assert np.abs(got[z * M:(z + 1) * M] - ref).max() < 2.0 ** -9 * max(np.abs(ref).max(), 1.0), f"slice {z}"


@pytest.mark.parametrize("M,N,K,batch,lda,ldb", [(1024, 64, 4096, 3, 1024, 64), (128, 128, 704, 5, 128, 128), (256, 64, 1000, 1, 256, 64),
                                                  (1024, 512, 2048, 2, 1024, 512), (200, 72, 320, 2, 264, 80), (512, 128, 96, 4, 512, 512)])
def test_gemm_tf32_tn_rows_contraction(cuda, M, N, K, batch, lda, ldb):
    """lpd_gemm_tf32_tn: C[z] = A[zK:(z+1)K]^T . B[zK:(z+1)K] with MN-major UMMA operands -> 2^-9 of the result scale"""
    r = rng(M + N + K + batch)
    A = r.standard_normal((batch * K, lda)).astype(np.float32)
    Bm = r.standard_normal((batch * K, ldb)).astype(np.float32)
    got = ops.gemm_tf32_tn(dev(A), dev(Bm), M=M, N=N, K=K, lda=lda, ldb=ldb, batch=batch).cpu().numpy()
    for z in range(batch):
        ref = A[z * K:(z + 1) * K, :M].astype(np.float64).T @ Bm[z * K:(z + 1) * K, :N].astype(np.float64)
        assert np.abs(got[z] - ref).max() < 2.0 ** -9 * np.abs(ref).max(), f"slice {z}"


@pytest.mark.parametrize("B,N,k,C", [(2, 300, 20, 128), (1, 257, 32, 128), (1, 100, 7, 128), (2, 200, 20, 64), (1, 90, 25, 64), (3, 1024, 20, 128)])
def test_edgeconv_dg_tensor_core_path(cuda, B, N, k, C):
    """lpd_edgeconv_dg_tf32: first layer exact fp32, second layer TF32 on tcgen05 -> 2^-9 of the layer-2 scale"""
    r = rng(N + k + C + 1)
    pq = r.standard_normal((B * N, 2 * C)).astype(np.float32)
    idx = r.integers(0, N, (B, N, k)).astype(np.int32)
    s1, t1, s2, t2 = (r.standard_normal(C).astype(np.float32) for _ in range(4))
    w2 = (r.standard_normal((C, C)) / np.sqrt(C)).astype(np.float32)
    P, Q = pq[:, :C].reshape(B, N, C), pq[:, C:].reshape(B, N, C)
    y1 = _leaky((P[np.arange(B)[:, None, None], idx] + Q[:, :, None, :]).astype(np.float64) * s1 + t1)
    y2 = _leaky((y1 @ w2.astype(np.float64).T) * s2 + t2)
    x1_ref, x2_ref = y1.max(2).reshape(B * N, C), y2.max(2).reshape(B * N, C)
    d_pq = dev(pq)
    x = torch.zeros(B * N, 2 * C, device="cuda")
    prev = ops.set_precision("tf32")
    try:
        ops.edgeconv_dg(d_pq, 2 * C, d_pq[:, C:], 2 * C, dev(idx, torch.int32), B, N, k, C, C, dev(s1), dev(t1), dev(w2), dev(s2), dev(t2),
                        ops.ACT_LEAKY, 0.01, x, 2 * C, x[:, C:], 2 * C)
    finally:
        ops.set_precision(prev)
    got = x.cpu().numpy()
    assert np.abs(got[:, :C] - x1_ref).max() < 1e-5
    scale2 = np.abs(y2).max()
    assert np.abs(got[:, C:] - x2_ref).max() < 2.0 ** -8 * scale2


@pytest.mark.parametrize("B,N", [(2, 300), (1, 4096), (3, 1001), (1, 23), (1, 20)])
def test_edgeconv_dg_prescaled_k20_kernel(cuda, B, N):
    """the specialised k = 20 / 128-channel kernel (edge_tc20.cu: 24-row point slots, two producer warps per point, first-layer
    BatchNorm pre-applied to p and q): x1 exact up to one fp32 add, x2 within 2^-8 of the layer-2 scale (TF32 second layer);
    point counts that do not fill the last 5-point tile, clouds smaller than a tile"""
    k, C = 20, 128
    r = rng(N + 77)
    pq = r.standard_normal((B * N, 2 * C)).astype(np.float32)
    idx = r.integers(0, N, (B, N, k)).astype(np.int32)
    s2, t2 = (r.standard_normal(C).astype(np.float32) for _ in range(2))
    w2 = (r.standard_normal((C, C)) / np.sqrt(C)).astype(np.float32)
    P, Q = pq[:, :C].reshape(B, N, C), pq[:, C:].reshape(B, N, C)
    y1 = _leaky((P[np.arange(B)[:, None, None], idx] + Q[:, :, None, :]).astype(np.float64))
    y2 = _leaky((y1 @ w2.astype(np.float64).T) * s2 + t2)
    x1_ref, x2_ref = y1.max(2).reshape(B * N, C), y2.max(2).reshape(B * N, C)
    d_pq = dev(pq)
    x = torch.zeros(B * N, 4 * C, device="cuda")
    prev = ops.set_precision("tf32")
    try:
        ops.profile(True)
        ops.edgeconv_dg(d_pq, 2 * C, d_pq[:, C:], 2 * C, dev(idx, torch.int32), B, N, k, C, C, None, None, dev(w2), dev(s2), dev(t2),
                        ops.ACT_LEAKY, 0.01, x, 4 * C, x[:, C:], 4 * C)
        labels = [l for l, _, _ in ops.profile(False)]
    finally:
        ops.set_precision(prev)
    assert labels == ["lpd_edgeconv_dg_tf32[128x128]"]
    got = x.cpu().numpy()
    assert np.abs(got[:, :C] - x1_ref).max() < 1e-6
    assert np.abs(got[:, C:2 * C] - x2_ref).max() < 2.0 ** -8 * np.abs(y2).max()
    assert (got[:, 2 * C:] == 0).all()
    with pytest.raises(Exception):      # the pre-scaled form exists for k == 20 only
        ops.set_precision("tf32")
        try:
            ops.edgeconv_dg(d_pq, 2 * C, d_pq[:, C:], 2 * C, dev(idx[:, :, :7].copy(), torch.int32), B, N, 7, C, C, None, None, dev(w2), dev(s2),
                            dev(t2), ops.ACT_LEAKY, 0.01, x, 4 * C, x[:, C:], 4 * C)
        finally:
            ops.set_precision(prev)


@pytest.mark.parametrize("B,N,k", [(2, 300, 20), (1, 1000, 32), (3, 64, 7)])
def test_edge_gather_max_f16(cuda, B, N, k):
    """fp16 rows, C = 256: out = fp16(act(q + max_m p_j)) — exact against numpy on the same fp16 inputs"""
    C, ld = 256, 512
    r = rng(N + k + 5)
    pq = r.standard_normal((B * N, ld)).astype(np.float16)
    idx = r.integers(0, N, (B, N, k)).astype(np.int32)
    P = pq[:, :C].reshape(B, N, C)
    mx = P[np.arange(B)[:, None, None], idx].max(2).reshape(B * N, C).astype(np.float32)
    v = mx + pq[:, C:].astype(np.float32)
    want = np.where(v > 0, v, v * np.float32(0.01)).astype(np.float16)
    d_pq = torch.from_numpy(pq).cuda()
    out = torch.zeros(B * N, ld, device="cuda", dtype=torch.float16)
    ops.edge_gather_max_f16(d_pq, ld, d_pq[:, C:], ld, dev(idx, torch.int32), B, N, k, C, ops.ACT_LEAKY, 0.01, out[:, C:], ld)
    got = out.cpu().numpy()
    assert np.array_equal(got[:, C:], want) and (got[:, :C] == 0).all()


@pytest.mark.parametrize("B,N,k", [(2, 300, 20), (1, 4096, 20), (3, 1001, 20), (1, 20, 20), (2, 300, 32), (1, 4099, 32), (1, 32, 32)])
def test_edgeconv_dg20_f16_kernel(cuda, B, N, k):
    """edge_tc20.cu, fp16 form: y1 = fp16(leaky(p_j + q_i)) in half2 arithmetic, second layer fp16 x fp16 -> fp32 on tcgen05,
    x1 / x2 written as fp16.  Against float64 on the same fp16 inputs: x1 within one fp16 rounding of the exact value, x2 within
    2^-9 of the layer-2 scale (fp16 rounding of y1, of the output, and of the slope constant on the negative branch).  k = 20 (five
    points per 128-edge tile) and k = 32 (four)."""
    C = 128
    r = rng(N + 78 + k)
    pq = r.standard_normal((B * N, 2 * C)).astype(np.float16)
    idx = r.integers(0, N, (B, N, k)).astype(np.int32)
    s2, t2 = (r.standard_normal(C).astype(np.float32) for _ in range(2))
    w2 = (r.standard_normal((C, C)) / np.sqrt(C)).astype(np.float16)
    P, Q = pq[:, :C].reshape(B, N, C).astype(np.float64), pq[:, C:].reshape(B, N, C).astype(np.float64)
    y1 = _leaky(P[np.arange(B)[:, None, None], idx] + Q[:, :, None, :])
    y2 = _leaky((y1 @ w2.astype(np.float64).T) * s2 + t2)
    x1_ref, x2_ref = y1.max(2).reshape(B * N, C), y2.max(2).reshape(B * N, C)
    d_pq = torch.from_numpy(pq).cuda()
    x = torch.zeros(B * N, 4 * C, device="cuda", dtype=torch.float16)
    ops.edgeconv_dg20_f16(d_pq, 2 * C, d_pq[:, C:], 2 * C, dev(idx, torch.int32), B, N, torch.from_numpy(w2).cuda(), dev(s2), dev(t2),
                          ops.ACT_LEAKY, 0.01, x, 4 * C, x[:, C:], 4 * C)
    got = x.float().cpu().numpy()
    assert np.abs(got[:, :C] - x1_ref).max() <= 2.0 ** -10 * max(1.0, np.abs(x1_ref).max())
    assert np.abs(got[:, C:2 * C] - x2_ref).max() <= 2.0 ** -9 * np.abs(y2).max()
    assert (got[:, 2 * C:] == 0).all()


# ------------------------------------------------------------------------------------------------ tensor-core kNN (exact)
@pytest.mark.parametrize("variant", [3, 2, 1, 0])
@pytest.mark.parametrize("B,N,k,kind", [
    (2, 512, 20, "relu"), (3, 130, 20, "relu"), (1, 2048, 32, "relu"), (2, 1000, 24, "relu"), (2, 1000, 25, "relu"),
    (1, 4096, 20, "relu"), (1, 300, 20, "dups"), (1, 256, 20, "same"), (1, 640, 20, "big"), (1, 128, 1, "relu"),
    (2, 700, 20, "clusters"), (2, 900, 20, "offset"), (1, 500, 20, "tiny"), (1, 500, 20, "huge"), (2, 1100, 32, "lattice"),
    (3, 257, 20, "relu"), (1, 129, 32, "relu"),
])
def test_knn_tensor_core_filter_refine_is_bit_exact(cuda, B, N, k, kind, variant):
    """lpd_knn_tc must return exactly the canonical neighbour lists for every filter formulation, including on inputs
    built to defeat the low-precision filter (duplicates, all-identical points, huge norms / offsets, tight clusters,
    lattices with masses of exact ties, extreme scales) where rows fall back to the exact kernel.  On well-behaved inputs
    the two-pass filter must resolve every tile itself (a filter that silently hands everything to the fallback would
    still be exact, but is not the kernel under test)."""
    r = rng(N + k)
    x = np.maximum(r.standard_normal((B, N, 64)), 0.01 * r.standard_normal((B, N, 64))).astype(np.float32)
    if kind == "dups":
        x[:, 100:200] = x[:, 0:100]
    elif kind == "same":
        x[:] = x[:, :1]
    elif kind == "big":
        x = (x * 300 + 1000).astype(np.float32)
    elif kind == "clusters":
        centres = r.standard_normal((B, 7, 64)).astype(np.float32) * 3
        x = (centres[:, r.integers(0, 7, N)] + 1e-3 * r.standard_normal((B, N, 64))).astype(np.float32)
    elif kind == "offset":
        x = (x + 40.0).astype(np.float32)
    elif kind == "tiny":
        x = (x * 1e-12).astype(np.float32)
    elif kind == "huge":
        x = (x * 1e14).astype(np.float32)
    elif kind == "lattice":
        x = r.integers(0, 3, (B, N, 64)).astype(np.float32)
        x[:, :, 8:] = 0
    want = knn_canonical(x, k)
    prev = ops.KNN_TENSOR_CORES
    prev_variant = ops.knn_tc_variant(variant)
    diag = {}
    try:
        ops.KNN_TENSOR_CORES = True
        ops.profile(True)
        got = ops.knn(dev(x), k, diag=diag).cpu().numpy()
        labels = [l for l, _, _ in ops.profile(False)]
        ops.KNN_TENSOR_CORES = False
        got_simt = ops.knn(dev(x), k).cpu().numpy()                    # CUDA-core kernel, same bits
    finally:
        ops.KNN_TENSOR_CORES = prev
        ops.knn_tc_variant(prev_variant)
    assert labels and labels[0].startswith("lpd_knn_tc")
    assert np.array_equal(got_simt, want)
    assert np.array_equal(got, want), f"{int((got != want).any(axis=2).sum())} rows differ; diag {diag}"
    # ("offset": the worst-case rounding bound of the canonical chain on uncentred features dominates there, the filter may
    # hand tiles to the exact kernel)
    if variant != 0 and kind in ("relu", "big", "tiny", "huge"):
        assert diag["flagged_tiles"] == 0, diag


def test_knn_tensor_core_variants_on_model_features(cuda):
    """C2-like input (real conv2 features of a spatially ordered cloud, N=4096, k=20): all three formulations agree bit for
    bit and the two-pass filter needs no fallback tile."""
    from lpdnet_b200 import synth
    from lpdnet_b200.util.PointNetVlad import PointNetVlad
    B, N, k = 3, 4096, 20
    model = PointNetVlad(num_points=N, featnet="lpdnet", emb_dims=1024)
    model.load_state_dict(synth.synthetic_state_dict(model))
    model = model.cuda().eval()
    emb = model.emb_nn
    p = emb._prep.get(emb, emb._build)
    with torch.no_grad():
        h, _, _, _ = emb._front(synth.clouds(B, N).cuda(), p, "LPDNet", True)
    feat = h.view(B, N, 64).contiguous()
    prev = ops.knn_tc_variant(-1)
    out = {}
    try:
        for v in (0, 1, 2, 3):
            ops.knn_tc_variant(v)
            diag = {}
            out[v] = (ops.knn(feat, k, diag=diag).cpu().numpy(), diag)
    finally:
        ops.knn_tc_variant(prev)
    want = knn_canonical(feat.cpu().numpy(), k)
    for v in (0, 1, 2, 3):
        assert np.array_equal(out[v][0], want), f"variant {v}"
    assert all(out[v][1]["flagged_tiles"] == 0 for v in (1, 2, 3)), [out[v][1] for v in (1, 2, 3)]
