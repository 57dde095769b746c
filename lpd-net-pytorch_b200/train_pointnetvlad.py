"""Drop-in for the hot-path part of the reference's train_pointnetvlad.py: run_model (:202-217, the tuple layout contract)
and the body of one training iteration (:121-130 / :150-159: zero_grad, run_model, loss, backward, optimizer.step).
Plus the checkpoint format and the learning-rate policy the reference wraps around it: save_model
(:172-199), the two load paths (:64-76) and ReduceLROnPlateau (:92).  The epoch loop, logging and the DataLoaders are control
plane and stay with the caller.
"""
from __future__ import annotations

import os

import torch

from .loss import pointnetvlad_loss as PNV_loss

__all__ = ["run_model", "train_step", "train_one_epoch", "train", "TrainConfig", "save_model", "load_checkpoint", "make_scheduler",
           "FEATURE_OUTPUT_DIM", "MODEL_FILENAME", "DIVISION_EPOCH"]

DIVISION_EPOCH = 7        # reference train_pointnetvlad.py:22: epochs <= 7 draw random negatives, later ones mined hard negatives
FEATURE_OUTPUT_DIM = 256  # reference config.py
MODEL_FILENAME = "model.ckpt"  # reference config.py:8


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


def save_model(model, optimizer, epoch, total_iterations, ave_one_percent_recall, model_save_path, best_so_far=None,
               filename=MODEL_FILENAME):
    """Reference :172-199.  Writes `<path>/<epoch>-model.ckpt` = {'epoch','iter','state_dict','optimizer','recall'} (the
    DataParallel wrapper, if any, is unwrapped) and, when the recall beats `best_so_far`, `<path>/best-model.ckpt`.
    Returns the updated best recall (the reference keeps it in a module global)."""
    model_to_save = model.module if isinstance(model, torch.nn.DataParallel) else model
    payload = {
        'epoch': epoch,
        'iter': total_iterations,
        'state_dict': model_to_save.state_dict(),
        'optimizer': optimizer.state_dict(),
        'recall': ave_one_percent_recall,
    }
    os.makedirs(model_save_path, exist_ok=True)
    torch.save(payload, os.path.join(model_save_path, str(epoch) + "-" + filename))
    best = -float("inf") if best_so_far is None else best_so_far
    if best < ave_one_percent_recall:
        best = ave_one_percent_recall
        torch.save(payload, os.path.join(model_save_path, "best" + "-" + filename))
    return best


def load_checkpoint(model, optimizer, pretrained_path, map_location=None):
    """Reference :64-76.  A path ending in '7' (`*.t7`) holds a bare state_dict and loads with strict=False; anything else
    is the dict written by save_model: strict load of 'state_dict', optimizer state restored.  Returns
    (starting_epoch, total_iterations); (0, 0) when the file does not exist or is a bare state_dict."""
    if not os.path.exists(pretrained_path):
        return 0, 0
    if pretrained_path[-1] == "7":
        model.load_state_dict(torch.load(pretrained_path, map_location=map_location), strict=False)
        return 0, 0
    checkpoint = torch.load(pretrained_path, map_location=map_location, weights_only=False)
    model.load_state_dict(checkpoint['state_dict'], strict=True)
    if optimizer is not None:
        optimizer.load_state_dict(checkpoint['optimizer'])
    return checkpoint['epoch'] + 1, checkpoint['iter']


def make_scheduler(optimizer):
    """Reference :92 (its `verbose=True` no longer exists in current torch): lr x 0.2 when the evaluation recall has not
    improved by 0.1 (relative) for 2 epochs, floor 1e-5; stepped with `scheduler.step(ave_one_percent_recall)`."""
    return torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, 'max', factor=0.2, patience=2, threshold=0.1, min_lr=0.00001)


# ---------------------------------------------------------------------------------------------------------------------
# Epoch loop (reference :38-170) with the reference's globals (para.args / para.model, TOTAL_ITERATIONS, the two DataLoaders,
# the tensorboard writer) turned into explicit arguments.  The schedule is the reference's: base loader up to DIVISION_EPOCH,
# then the hard-negative loader with a descriptor refresh at the switch and every 700 (epoch + 1) samples; evaluation, checkpoint
# and ReduceLROnPlateau('max') after every epoch.
# ---------------------------------------------------------------------------------------------------------------------
class TrainConfig:
    """the argparse flags of util/initPara.py the loop reads, with the reference's defaults (:29-90)"""

    def __init__(self, batch_num_queries=2, max_epoch=20, lr=1e-3, optimizer="adam", momentum=0.9, loss_function="quadruplet",
                 margin_1=0.5, margin_2=0.2, triplet_use_best_positives=True, loss_lazy=True, loss_ignore_zero_batch=False,
                 pretrained_path="", model_save_path="checkpoints"):
        self.__dict__.update(locals())
        del self.__dict__["self"]


def train_one_epoch(model, optimizer, loss_function, epoch, loader_base, loader_advance, cfg: TrainConfig, state: dict,
                    update_vectors=None, log=None):
    """Reference :117-170.  `loader_base` / `loader_advance` yield (queries, positives, negatives, other_neg) batches in the
    DataLoader layout; `update_vectors()` re-embeds the training set for hard-negative mining (util/data.py:277-354);
    `state["iter"]` is the reference's TOTAL_ITERATIONS; `log(name, value, iteration)` replaces the tensorboard writer."""
    batch_num = cfg.batch_num_queries
    advance = epoch > DIVISION_EPOCH
    if advance and epoch == DIVISION_EPOCH + 1 and update_vectors is not None:
        update_vectors()
    for queries, positives, negatives, other_neg in (loader_advance if advance else loader_base):
        loss = train_step(model, optimizer, queries, positives, negatives, other_neg, cfg.margin_1, cfg.margin_2,
                          loss_function=loss_function, use_min=cfg.triplet_use_best_positives, lazy=cfg.loss_lazy,
                          ignore_zero_loss=cfg.loss_ignore_zero_batch)
        if log is not None:
            log("epoch", epoch, state["iter"])
            log("Loss", loss.cpu().item(), state["iter"])
            log("learn rate", optimizer.param_groups[0]["lr"], state["iter"])
        state["iter"] += batch_num
        if advance and update_vectors is not None:
            period = int(700 * (epoch + 1)) // batch_num * batch_num
            if period > 0 and state["iter"] % period == 0:
                update_vectors()
    return state


def train(model, loader_base, loader_advance, evaluate_fn, cfg: TrainConfig | None = None, update_vectors=None, log=None):
    """Reference :38-115.  `evaluate_fn(model) -> (ave_recall, average_similarity, ave_one_percent_recall)` is
    evaluate.evaluate_model bound to the evaluation sets.  Returns the state dict {"epoch", "iter", "best", "recall"}.
    With torch.distributed initialised every rank runs this loop on its own tuples (lpdnet_b200.optim.Adam reduces the
    gradients); only rank 0 writes checkpoints."""
    import torch.distributed as dist
    from . import optim as lpd_optim
    cfg = cfg or TrainConfig()
    loss_function = PNV_loss.quadruplet_loss if cfg.loss_function == "quadruplet" else PNV_loss.triplet_loss_wrapper   # :43-49
    if cfg.optimizer == "momentum":                                                                                    # :51-61
        optimizer = torch.optim.SGD(model.parameters(), cfg.lr, momentum=cfg.momentum)
    elif cfg.optimizer == "adam":
        optimizer = lpd_optim.Adam(model.parameters(), cfg.lr)
    else:
        raise ValueError(f"optimizer {cfg.optimizer!r}: the reference knows 'adam' and 'momentum'")
    starting_epoch, total_iterations = (0, 0)
    if cfg.pretrained_path:
        starting_epoch, total_iterations = load_checkpoint(model, optimizer, cfg.pretrained_path)                    # :64-77
    state = {"epoch": starting_epoch, "iter": total_iterations, "best": None, "recall": 0}
    if starting_epoch > DIVISION_EPOCH + 1 and update_vectors is not None:                                             # :86-87
        update_vectors()
    scheduler = make_scheduler(optimizer)                                                                              # :92
    rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
    for epoch in range(starting_epoch, cfg.max_epoch):                                                                 # :101-115
        train_one_epoch(model, optimizer, loss_function, epoch, loader_base, loader_advance, cfg, state, update_vectors, log)
        _, _, ave_one_percent_recall = evaluate_fn(model)
        state["epoch"], state["recall"] = epoch, ave_one_percent_recall
        if rank0:
            state["best"] = save_model(model, optimizer, epoch, state["iter"], ave_one_percent_recall, cfg.model_save_path,
                                       state["best"])
        scheduler.step(ave_one_percent_recall)
        if log is not None:
            log("Val Recall", ave_one_percent_recall, epoch)
    return state
