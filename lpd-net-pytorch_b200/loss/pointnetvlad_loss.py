"""Drop-in for the reference's loss/pointnetvlad_loss.py (best_pos_distance :6-12, triplet_loss :15-42,
triplet_loss_wrapper :45-46, quadruplet_loss :49-97): same signatures and argument meaning.

Each loss is ONE kernel launch (lpd_quadruplet_loss) that produces the scalar and, when any input requires
grad, all four input gradients; backward() only scales them by the incoming gradient.
"""
from __future__ import annotations

import torch

from .. import ops
from .._host import require_cuda

__all__ = ["best_pos_distance", "triplet_loss", "triplet_loss_wrapper", "quadruplet_loss"]


class _HingeLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_vec, pos_vecs, neg_vecs, other_neg, m1, m2, use_min, lazy, ignore_zero_loss):
        need = any(t is not None and t.requires_grad for t in (q_vec, pos_vecs, neg_vecs, other_neg))
        if need:
            loss, grads = ops.quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2, use_min, lazy,
                                              ignore_zero_loss, need_grad=True)
            ctx.save_for_backward(*[g for g in grads if g is not None])
            ctx.has_other = other_neg is not None
        else:
            loss = ops.quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2, use_min, lazy, ignore_zero_loss)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        gq, gpos, gneg = (g * grad_out for g in saved[:3])
        gother = saved[3] * grad_out if ctx.has_other else None
        return gq, gpos, gneg, gother, None, None, None, None, None


def _check(*ts):
    for t in ts:
        if t is not None:
            require_cuda(t, "pointnetvlad_loss")


def best_pos_distance(query, pos_vecs):
    """(min_p, max_p) of ||pos_p - query||^2 per query tuple: [Bq], [Bq] (reference :6-12)."""
    _check(query, pos_vecs)
    Bq, P, D = pos_vecs.shape
    if P > 32:
        raise NotImplementedError("best_pos_distance: at most 32 positives per query")
    # one launch for all tuples: the positives of tuple b are database segment b of the exact per-segment search; the sorted [P]
    # distance table of (query b, segment b) holds the minimum first and the maximum last
    db = pos_vecs.detach().reshape(Bq * P, D)
    q = query.detach().reshape(Bq, D)
    if D % 4 == 0:
        seg = torch.arange(0, (Bq + 1) * P, P, dtype=torch.int32, device=db.device)
        _, dist = ops.retrieval_tc(db, q, P, seg)                                  # [segment, query, P], ascending
        d = dist[torch.arange(Bq, device=db.device), torch.arange(Bq, device=db.device)]
        return d[:, 0].float(), d[:, P - 1].float()
    mins, maxs = [], []
    for b in range(Bq):          # (descriptor sizes the tensor-core search does not take: one exact search per tuple)
        _, d = ops.retrieval_topk(db[b * P:(b + 1) * P], q[b:b + 1], P)
        mins.append(d[0, 0])
        maxs.append(d[0, P - 1])
    return torch.stack(mins).float(), torch.stack(maxs).float()


def triplet_loss(q_vec, pos_vecs, neg_vecs, margin, use_min=False, lazy=False, ignore_zero_loss=False):
    _check(q_vec, pos_vecs, neg_vecs)
    return _HingeLoss.apply(q_vec, pos_vecs, neg_vecs, None, float(margin), 0.0, bool(use_min), bool(lazy),
                            bool(ignore_zero_loss))


def triplet_loss_wrapper(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2, use_min=False, lazy=False, ignore_zero_loss=False):
    return triplet_loss(q_vec, pos_vecs, neg_vecs, m1, use_min, lazy, ignore_zero_loss)


def quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2, use_min=False, lazy=False, ignore_zero_loss=False):
    _check(q_vec, pos_vecs, neg_vecs, other_neg)
    return _HingeLoss.apply(q_vec, pos_vecs, neg_vecs, other_neg, float(m1), float(m2), bool(use_min), bool(lazy),
                            bool(ignore_zero_loss))
