"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  numpy restatement of reference loss/pointnetvlad_loss.py.

  best_pos_distance  :6-12     triplet_loss :15-42     triplet_loss_wrapper :45-46     quadruplet_loss :49-97
Shapes: q [Bq,1,D], pos [Bq,P,D], neg [Bq,Nn,D], other [Bq,1,D]; float32 arithmetic like the reference.
`quadruplet_loss_grad` returns the autograd gradients (min/max route to the first extremal element,
clamp(min=0) passes the gradient where its input is >= 0).
"""
import numpy as np

f32 = np.float32


def best_pos_distance(query, pos_vecs):
    diff = ((pos_vecs - query) ** 2).sum(2, dtype=f32)
    return diff.min(1), diff.max(1)


def _hinge(positive, neg_vecs, anchor, margin, lazy, ignore_zero_loss):
    loss = f32(margin) + positive[:, None] - ((neg_vecs - anchor) ** 2).sum(2, dtype=f32)
    loss = np.maximum(loss, f32(0))
    t = loss.max(1) if lazy else loss.sum(1, dtype=f32)
    if ignore_zero_loss:
        hard = (t > 1e-16).astype(f32).sum()
        return f32(t.sum(dtype=f32) / (hard + f32(1e-16)))
    return f32(t.mean(dtype=f32))


def triplet_loss(q_vec, pos_vecs, neg_vecs, margin, use_min=False, lazy=False, ignore_zero_loss=False):
    mn, mx = best_pos_distance(q_vec, pos_vecs)
    return _hinge(mn if use_min else mx, neg_vecs, q_vec, margin, lazy, ignore_zero_loss)


def triplet_loss_wrapper(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2, use_min=False, lazy=False, ignore_zero_loss=False):
    return triplet_loss(q_vec, pos_vecs, neg_vecs, m1, use_min, lazy, ignore_zero_loss)


def quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2, use_min=False, lazy=False, ignore_zero_loss=False):
    mn, mx = best_pos_distance(q_vec, pos_vecs)
    positive = mn if use_min else mx
    return f32(_hinge(positive, neg_vecs, q_vec, m1, lazy, ignore_zero_loss)
               + _hinge(positive, neg_vecs, other_neg, m2, lazy, ignore_zero_loss))


def quadruplet_loss_grad(q, pos, neg, other, m1, m2, use_min=False, lazy=False, ignore_zero_loss=False):
    """Gradients of quadruplet_loss (other=None: triplet) in float64, for gradient-parity checks."""
    q, pos, neg = (np.asarray(a, dtype=np.float64) for a in (q, pos, neg))
    oth = None if other is None else np.asarray(other, dtype=np.float64)
    Bq, P, D = pos.shape
    Nn = neg.shape[1]
    dpos = ((pos - q) ** 2).sum(2)
    ps = dpos.argmin(1) if use_min else dpos.argmax(1)
    positive = dpos[np.arange(Bq), ps]
    gq, gpos, gneg = np.zeros_like(q), np.zeros_like(pos), np.zeros_like(neg)
    goth = None if oth is None else np.zeros_like(oth)
    for anchor, margin, ganchor in ((q, m1, gq), (oth, m2, goth)):
        if anchor is None:
            continue
        pre = margin + positive[:, None] - ((neg - anchor) ** 2).sum(2)
        l = np.maximum(pre, 0)
        t = l.max(1) if lazy else l.sum(1)
        coef = 1.0 / ((t > 1e-16).sum() + 1e-16) if ignore_zero_loss else 1.0 / Bq
        mask = np.zeros_like(l)
        if lazy:
            mask[np.arange(Bq), l.argmax(1)] = 1.0
        else:
            mask[:] = 1.0
        a = coef * mask * (pre >= 0)                      # dL/dl
        diff = 2.0 * (neg - anchor)                       # d ||neg-anchor||^2 / d neg
        gneg -= a[:, :, None] * diff
        ganchor += (a[:, :, None] * diff).sum(1, keepdims=True)
        ap = a.sum(1)                                     # dL/dpositive
        dp = 2.0 * (pos[np.arange(Bq), ps] - q[:, 0])
        gpos[np.arange(Bq), ps] += ap[:, None] * dp
        gq[:, 0] -= ap[:, None] * dp
    return gq, gpos, gneg, goth
