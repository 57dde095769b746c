"""Shared test helpers (tests/ is on sys.path under pytest's default import mode)."""
import numpy as np


def canonical_pd(x_pm):
    """fp64 -||xi-xj||^2 for the tie-aware comparison"""
    x = x_pm.astype(np.float64)
    g = x @ x.T
    n = (x * x).sum(1)
    return -(n[:, None] + n[None, :] - 2 * g)


def assert_knn_equivalent(x_pm, idx_a, idx_b, k, ulps=8):
    """Rows must hold the same neighbour SET, or differ only in candidates whose distance is within a few
    ulp of the k-th distance (the reference's SGEMM order / topk tie order are unspecified)."""
    B = x_pm.shape[0]
    bad = 0
    for b in range(B):
        pd = None
        for i in range(x_pm.shape[1]):
            sa, sb = set(idx_a[b, i].tolist()), set(idx_b[b, i].tolist())
            if sa == sb:
                continue
            if pd is None:
                pd = canonical_pd(x_pm[b])
            kth = np.sort(pd[i])[::-1][k - 1]
            scale = np.abs(x_pm[b]).max() ** 2 * x_pm.shape[2]
            tol = ulps * np.finfo(np.float32).eps * max(scale, 1e-30)
            for j in sa ^ sb:
                assert abs(pd[i, j] - kth) <= tol, f"cloud {b} row {i}: contested neighbour {j} is not a near-tie"
            bad += 1
    return bad
