"""Drop-in for the hot-path callers in the reference's util/data.py:

    get_random_hard_negatives   :103-115   KDTree over <= a few thousand cached descriptors per sample -> hardest negatives
    get_feature_representation  :117-133   one submap -> descriptor, eval mode, model.train() restored
    update_vectors              :277-354   bulk eval-mode embedding of the training set into TRAINING_LATENT_VECTORS

The KDTree query becomes ONE exact brute-force top-k launch (lpd_retrieval_topk: fp64 distances like sklearn, ties to the
lower index); `hard_negatives_batch` mines a whole batch of samples in one launch per sample group.  The embedding goes
through the pinned, double-buffered driver of lpdnet_b200.evaluate.  There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import evaluate, ops

TRAINING_LATENT_VECTORS = []  # reference util/data.py global (filled by update_vectors)


def _dev():
    return evaluate._device()


def get_random_hard_negatives(query_vec, random_negs, hard_neg_num, latent_vectors=None):
    """Same contract as the reference (:103-115): indices (entries of `random_negs`) of the `hard_neg_num` cached descriptors
    closest to `query_vec`, nearest first.  `latent_vectors` defaults to the module global the reference uses."""
    table = TRAINING_LATENT_VECTORS if latent_vectors is None else latent_vectors
    negs = np.asarray(random_negs)
    if isinstance(table, torch.Tensor):
        latent = table[torch.as_tensor(negs, device=table.device, dtype=torch.long)].to(_dev(), torch.float32)
    else:
        latent = torch.from_numpy(np.ascontiguousarray(np.asarray(table)[negs], dtype=np.float32)).to(_dev())
    q = torch.as_tensor(np.asarray(query_vec, dtype=np.float32)).reshape(1, -1).to(_dev())
    idx, _ = ops.retrieval_topk(latent.contiguous(), q.contiguous(), int(hard_neg_num), want_dist=False)
    hard_negs = np.squeeze(negs[idx[0].cpu().numpy()])
    return hard_negs.tolist()


def hard_negatives_batch(query_vecs, latent_vectors, neg_lists, hard_neg_num):
    """Batched form for a DataLoader-free pipeline: `neg_lists[i]` are the candidate negatives of query i.  One gather + one
    launch per query keeps the exact KDTree semantics (each query has its own candidate set) while the descriptor table
    stays resident on the device."""
    dev = _dev()
    table = latent_vectors if isinstance(latent_vectors, torch.Tensor) else torch.from_numpy(
        np.ascontiguousarray(latent_vectors, dtype=np.float32))
    table = table.to(dev, torch.float32)
    q = torch.as_tensor(np.asarray(query_vecs, dtype=np.float32)).to(dev)
    out = []
    for i, negs in enumerate(neg_lists):
        negs_t = torch.as_tensor(np.asarray(negs), device=dev, dtype=torch.long)
        idx, _ = ops.retrieval_topk(table[negs_t].contiguous(), q[i:i + 1].contiguous(), int(hard_neg_num), want_dist=False)
        out.append(negs_t[idx[0].long()].cpu().tolist())
    return out


def get_feature_representation(cloud, model):
    """reference :117-133 over an in-memory submap ([N, 3] array; the file read is loading_pointclouds.load_pc_files):
    eval-mode descriptor [256] as numpy, then model.train() exactly like the reference."""
    model.eval()
    q = torch.from_numpy(np.ascontiguousarray(np.asarray(cloud), dtype=np.float32)).reshape(1, 1, -1, 3).to(_dev())
    with torch.no_grad():
        output = model(q)
    output = np.squeeze(output.detach().cpu().numpy())
    model.train()
    return output


def update_vectors(model, clouds, batch_num: int = 64):
    """reference :277-354 re-expressed over the in-RAM training cloud cache ([n, N, 3], the reference's
    TRAINING_POINT_CLOUD): embeds every training submap in eval mode and replaces TRAINING_LATENT_VECTORS.  The model is
    left in train mode, as the reference does (:352)."""
    global TRAINING_LATENT_VECTORS
    vecs = evaluate.get_latent_vectors(model, clouds, batch_num=batch_num)
    model.train()
    TRAINING_LATENT_VECTORS = vecs
    return vecs
