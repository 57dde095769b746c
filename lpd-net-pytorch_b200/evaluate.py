"""Drop-in for the hot-path part of the reference's evaluate.py: get_recall (:162-206) with the same signature
and return triple, and the bulk embedding driver get_latent_vectors (:96-159) re-expressed over in-memory clouds.

get_recall replaces the per-query sklearn KDTree.query loop by ONE exact brute-force top-25 kernel launch per
(database run, query run) pair (lpd_retrieval_topk: fp64 distances like KDTree, ties to the lower index); the
counting logic after the search is the reference's, kept on the host because it consumes Python ground-truth
lists (QUERY_SETS[n][i][m]).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

__all__ = ["get_recall", "get_latent_vectors", "evaluate_sets", "recall_num"]

recall_num = 25  # reference evaluate.py:20


def _device():
    if not torch.cuda.is_available():
        from ._lib import LpdError
        raise LpdError("evaluate: no CUDA device; the retrieval kernel has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _as_dev(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.detach().to(_device(), torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(_device())


def get_recall(m, n, DATABASE_VECTORS, QUERY_VECTORS, QUERY_SETS):
    """Same contract as the reference: returns (recall[25] cumulative %, top1_similarity_score list, one_percent_recall %)."""
    database_output = DATABASE_VECTORS[m]
    queries_output = QUERY_VECTORS[n]
    db_dev, q_dev = _as_dev(database_output), _as_dev(queries_output)
    k = min(recall_num, db_dev.shape[0])
    idx, _ = ops.retrieval_topk(db_dev, q_dev, k, want_dist=False)
    indices = idx.cpu().numpy()

    recall = [0] * recall_num
    top1_similarity_score = []
    one_percent_retrieved = 0
    threshold = max(int(round(len(database_output) / 100.0)), 1)
    num_evaluated = 0
    db_host = database_output.detach().cpu().numpy() if isinstance(database_output, torch.Tensor) else np.asarray(database_output)
    q_host = queries_output.detach().cpu().numpy() if isinstance(queries_output, torch.Tensor) else np.asarray(queries_output)
    for i in range(len(q_host)):
        true_neighbors = QUERY_SETS[n][i][m]
        if len(true_neighbors) == 0:
            continue
        num_evaluated += 1
        row = indices[i]
        truth = set(true_neighbors)
        for j in range(len(row)):
            if row[j] in truth:
                if j == 0:
                    top1_similarity_score.append(np.dot(q_host[i], db_host[row[j]]))
                recall[j] += 1
                break
        if len(set(row[0:threshold].tolist()).intersection(truth)) > 0:
            one_percent_retrieved += 1
    one_percent_recall = (one_percent_retrieved / float(num_evaluated)) * 100
    recall = (np.cumsum(recall) / float(num_evaluated)) * 100
    return recall, top1_similarity_score, one_percent_recall


def get_latent_vectors(model, clouds, batch_num: int = 64, pin: bool = True) -> np.ndarray:
    """Eval-mode embedding of `clouds` ([n, N, 3] numpy / tensor, host memory) in batches of `batch_num`
    (reference :96-159: batch = eval_batch_size * (1 + P + Nn), tail batch handled, model.eval()/train() toggled).
    Host->device copies are issued from pinned memory on a side stream so they overlap the previous batch."""
    was_training = model.training
    model.eval()
    dev = _device()
    x = clouds if isinstance(clouds, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(clouds, dtype=np.float32))
    x = x.float()
    if x.dim() == 3:
        x = x.unsqueeze(1)
    n = x.shape[0]
    out = torch.empty(n, model.net_vlad.output_dim if hasattr(model, "net_vlad") else 256, dtype=torch.float32,
                      pin_memory=pin)
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    staged = None

    def stage(lo):
        hi = min(n, lo + batch_num)
        src = x[lo:hi].pin_memory() if pin and not x.is_pinned() else x[lo:hi]
        with torch.cuda.stream(copy_stream):
            d = src.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return lo, hi, d, ev, src

    with torch.no_grad():
        if n:
            staged = stage(0)
        while staged is not None:
            lo, hi, d, ev, _keep = staged
            staged = stage(hi) if hi < n else None
            main.wait_event(ev)
            d.record_stream(main)
            o = model(d)
            out[lo:hi].copy_(o, non_blocking=True)
    torch.cuda.synchronize(dev)
    model.train(was_training)
    return out.numpy()


def evaluate_sets(DATABASE_VECTORS, QUERY_VECTORS, QUERY_SETS):
    """Pair loop + aggregation of evaluate_model (reference :59-93) over already-embedded sets:
    returns (ave_recall[25], average_similarity, ave_one_percent_recall)."""
    recall = np.zeros(recall_num)
    count = 0
    similarity = []
    one_percent_recall = []
    for m in range(len(QUERY_SETS)):
        for n in range(len(QUERY_SETS)):
            if m == n:
                continue
            pair_recall, pair_similarity, pair_opr = get_recall(m, n, DATABASE_VECTORS, QUERY_VECTORS, QUERY_SETS)
            recall += np.array(pair_recall)
            count += 1
            one_percent_recall.append(pair_opr)
            for x in pair_similarity:
                similarity.append(x)
    ave_recall = recall / count
    average_similarity = np.mean(similarity)
    ave_one_percent_recall = np.mean(one_percent_recall)
    return ave_recall, average_similarity, ave_one_percent_recall
