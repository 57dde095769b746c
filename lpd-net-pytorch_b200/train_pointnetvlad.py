"""Drop-in for the hot-path part of the reference's train_pointnetvlad.py: run_model (:202-217, the tuple layout contract)
and the body of one training iteration (:121-130 / :150-159: zero_grad, run_model, loss, backward, optimizer.step).
The epoch loop, logging, checkpoint naming and the DataLoaders around it are control plane and stay with the caller.
"""
from __future__ import annotations

import torch

from .loss import pointnetvlad_loss as PNV_loss

__all__ = ["run_model", "train_step", "FEATURE_OUTPUT_DIM"]

FEATURE_OUTPUT_DIM = 256  # reference config.py


def run_model(model, queries, positives, negatives, other_neg, require_grad=True, num_points=None,
              feature_output_dim=FEATURE_OUTPUT_DIM):
    """Reference :202-217.  queries [Bq,1,N,3], positives [Bq,P,N,3], negatives [Bq,Nn,N,3], other_neg [Bq,1,N,3]
    (host or device) -> (o_q [Bq,1,D], o_pos [Bq,P,D], o_neg [Bq,Nn,D], o_other [Bq,1,D]).  The tuple members of one
    query are contiguous in the fed batch: cat on dim 1, then view(-1, 1, N, 3)."""
    Bq, P, Nn = queries.shape[0], positives.shape[1], negatives.shape[1]
    N = queries.shape[-2] if num_points is None else num_points
    feed = torch.cat((queries, positives, negatives, other_neg), 1).view(-1, 1, N, 3)
    dev = next(model.parameters()).device
    feed = feed.to(dev, non_blocking=True).float()
    if require_grad:
        output = model(feed)
    else:
        with torch.no_grad():
            output = model(feed)
    output = output.view(Bq, -1, feature_output_dim)
    return torch.split(output, [1, P, Nn, 1], dim=1)


def train_step(model, optimizer, queries, positives, negatives, other_neg, margin_1=0.5, margin_2=0.2,
               loss_function=PNV_loss.quadruplet_loss, use_min=True, lazy=True, ignore_zero_loss=False):
    """One iteration of the reference's train_one_epoch body (:121-130).  Returns the loss tensor (device scalar).
    With torch.distributed initialised and lpdnet_b200.optim.Adam, optimizer.step() all-reduces the flat gradient
    buffer over NCCL first (each rank owns whole tuples)."""
    model.train()
    optimizer.zero_grad()
    o_q, o_pos, o_neg, o_other = run_model(model, queries, positives, negatives, other_neg)
    loss = loss_function(o_q, o_pos, o_neg, o_other, margin_1, margin_2, use_min=use_min, lazy=lazy,
                         ignore_zero_loss=ignore_zero_loss)
    loss.backward()
    optimizer.step()
    return loss.detach()
