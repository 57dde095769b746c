"""CPU check of the ARGUMENT behind the two-pass tensor-core kNN filter (lpd-net-pytorch_b200/csrc/knn_tc2.cu), independent of
the GPU: a numpy emulation of the filter (centring, power-of-two scaling, fp16 operands, strided groups under the per-block
affine scramble, tau0 = k-th largest group maximum, collection of everything >= tau0 - 2 eps with the kernel's eps formula)
must always collect a superset of the canonical top-k of oracle/knn_canonical.c - on benign and on adversarial clouds - and
stay selective on benign ones."""
import numpy as np
import pytest

from oracle import knn_canonical


def block_perm(t):
    v = ((t + 1) * 2654435761) & 0xFFFFFFFF
    return ((v >> 8) & 63) | 1, (v >> 20) & 63


def emulate_filter(x, k, tight):
    """x [N, 64] float32 -> (collected boolean [N, N], eps [N]); mirrors knn2_center / knn2_prep / knn2_tc_kernel"""
    N = x.shape[0]
    mu = (x.astype(np.float32).sum(0, dtype=np.float32) / np.float32(N)).astype(np.float32)
    xc = (x - mu).astype(np.float32)
    g = float(np.abs(xc).max())
    se = 8 - (int(np.frexp(np.float32(g))[1]) if g > 0 else 8)
    sigma = np.float32(2.0 ** se)
    xh = (xc * sigma).astype(np.float16)                                   # round to nearest even, like __floats2half2_rn
    nrm = (xc.astype(np.float32) ** 2).sum(1, dtype=np.float32)
    xx = (x.astype(np.float32) ** 2).sum(1, dtype=np.float32)
    inv_cb = np.float64(sigma) ** 2 / 2.0
    t = xh.astype(np.float64) @ xh.astype(np.float64).T - (nrm.astype(np.float64) * inv_cb)[None, :]   # tensor-core scores
    ni, rc = np.sqrt(nrm.astype(np.float64)), np.sqrt(float(nrm.max()))
    eps = (1.9e-6 / float(sigma)) * (ni + rc) + 2.0 ** -17 * (xx.astype(np.float64) + float(xx.max())) + 2.0 ** -18 * rc * rc
    sn = np.sqrt(nrm.astype(np.float64))
    if tight:
        e = 2.5e-3 * inv_cb * ni[:, None] * sn[None, :]
        lower, upper = t - e, t + e
    else:
        eps = eps + 2.5e-3 * ni * rc
        lower = upper = t
    # strided groups: storage position inside the (scrambled) 64-block
    grp = np.arange(N) & 63
    for blk in range(N >> 6):
        a, b = block_perm(blk)
        grp[blk * 64:(blk + 1) * 64] = (np.arange(64) * a + b) & 63
    gmax = np.full((N, 64), -np.inf)
    for gi in range(64):
        sel = grp == gi
        if sel.any():
            gmax[:, gi] = lower[:, sel].max(1)
    tau0 = np.sort(gmax, axis=1)[:, ::-1][:, k - 1]
    thr = np.maximum(tau0 - 2.0 * eps * inv_cb, -1.0e9)
    return upper >= thr[:, None], eps


def clouds(kind, N, r):
    x = np.maximum(r.standard_normal((N, 64)), 0.01 * r.standard_normal((N, 64))).astype(np.float32)
    if kind == "clusters":
        centres = r.standard_normal((7, 64)).astype(np.float32) * 3
        x = (centres[r.integers(0, 7, N)] + 1e-3 * r.standard_normal((N, 64))).astype(np.float32)
    elif kind == "offset":
        x = (x + 40.0).astype(np.float32)
    elif kind == "big":
        x = (x * 300 + 1000).astype(np.float32)
    elif kind == "tiny":
        x = (x * 1e-12).astype(np.float32)
    elif kind == "manifold":                                                # smooth 3-d manifold in 64-d, like real conv2 features
        p = r.uniform(-1, 1, (N, 3)).astype(np.float32)
        w1, w2 = r.standard_normal((3, 64)).astype(np.float32), (r.standard_normal((64, 64)) / 8).astype(np.float32)
        h = np.maximum(p @ w1, 0.01 * (p @ w1))
        x = np.maximum(h @ w2, 0.01 * (h @ w2)).astype(np.float32)
    elif kind == "dups":
        x[100:200] = x[0:100]
    return x


@pytest.mark.parametrize("kind,N,k", [("relu", 1024, 20), ("manifold", 2048, 20), ("manifold", 1536, 32), ("clusters", 700, 20),
                                      ("offset", 900, 20), ("big", 640, 20), ("tiny", 500, 20), ("dups", 512, 20), ("relu", 300, 25)])
def test_two_pass_filter_collects_a_superset_of_the_canonical_topk(kind, N, k):
    r = np.random.default_rng(N + k)
    x = clouds(kind, N, r)
    want = knn_canonical(x[None], k)[0]
    collected, _ = emulate_filter(x, k, tight=k > 24)
    rows = np.arange(N)[:, None]
    assert collected[rows, want].all(), f"{int((~collected[rows, want]).any(1).sum())} rows would lose a canonical neighbour"
    if kind in ("relu", "manifold"):                                        # and the filter is selective where it should be
        per_row = collected.sum(1)
        assert per_row.mean() < 2.2 * k and per_row.max() <= 2 * (40 if k <= 24 else 64), (per_row.mean(), per_row.max())
