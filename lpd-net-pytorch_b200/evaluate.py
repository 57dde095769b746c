"""Drop-in for the hot-path part of the reference's evaluate.py: get_recall (:162-206) with the same signature and return
triple, the pair loop + aggregation of evaluate_model (:59-93), and the bulk embedding driver get_latent_vectors (:96-159)
re-expressed over in-memory clouds.

Retrieval replaces the per-query sklearn KDTree.query loop: the databases of all runs are stacked and searched in ONE pass
(lpd_retrieval_tc: 3xTF32 distance GEMM on the tensor cores as a filter, exact fp64 difference-form re-rank like KDTree, ties
to the lower index), and the first-hit histogram / recall@1% / top-1 similarity bookkeeping of get_recall runs on the device
as well (lpd_recall_count).  The Python ground-truth lists (QUERY_SETS[n][i][m]) are flattened once into a CSR table
(`prepare_truth`); no per-query Python loop remains.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

__all__ = ["get_recall", "get_latent_vectors", "evaluate_sets", "evaluate_model", "recall_all_pairs", "prepare_truth", "recall_num"]

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


def _threshold(n_db: int) -> int:
    return max(int(round(n_db / 100.0)), 1)          # reference :174 (Python's round: half to even)


class Truth:
    """QUERY_SETS flattened for the device: query i of run n is row q_off[n] + i of the stacked query matrix; its true
    neighbours in database run m are truth_idx[truth_off[row * R + m] : truth_off[row * R + m + 1]] (indices local to run m)."""

    def __init__(self, q_sizes, truth_off, truth_idx):
        self.q_sizes = list(q_sizes)
        self.R = len(self.q_sizes)
        self.q_off = np.concatenate(([0], np.cumsum(self.q_sizes))).astype(np.int64)
        self.q_run = np.repeat(np.arange(self.R, dtype=np.int32), self.q_sizes)
        self.truth_off = truth_off
        self.truth_idx = truth_idx
        self._dev = None

    def device(self, dev):
        if self._dev is None or self._dev[0] != dev:
            self._dev = (dev, torch.from_numpy(self.q_run).to(dev), torch.from_numpy(self.truth_off).to(dev),
                         torch.from_numpy(self.truth_idx if self.truth_idx.size else np.zeros(1, np.int32)).to(dev))
        return self._dev[1:]


def prepare_truth(QUERY_SETS, runs=None) -> Truth:
    """One pass over the reference's ground-truth structure (list over runs n of per-query dicts whose key m holds the list of
    true database indices in run m, generating_queries/generate_test_sets.py:86-112) -> CSR arrays.  Host-side data
    preparation, done once per evaluation set."""
    R = len(QUERY_SETS) if runs is None else runs
    flat, offs, sizes = [], [0], []
    for n in range(R):
        qs = QUERY_SETS[n]
        sizes.append(len(qs))
        for i in range(len(qs)):
            entry = qs[i]
            for m in range(R):
                t = entry[m]
                if len(t):
                    flat.extend(t)
                offs.append(len(flat))
    return Truth(sizes, np.asarray(offs, dtype=np.int32), np.asarray(flat, dtype=np.int32))


def recall_all_pairs(DATABASE_VECTORS, QUERY_VECTORS, QUERY_SETS, truth: Truth | None = None, group=None):
    """Every ordered (database run m, query run n) pair of the evaluation in one pass.  Returns a dict:
         recall   [R, R, 25]  cumulative recall@1..25 in percent, indexed [n][m]   (get_recall's first return value)
         one_pct  [R, R]      recall@1% in percent                               (third return value)
         sim      [Nq, R]     float32 top-1 similarity where the top-1 is a true neighbour, NaN elsewhere (second)
         n_eval   [R, R]      queries with non-empty ground truth
       With torch.distributed initialised the database RUNS are sharded over the ranks (each rank searches all queries in its
       runs' segments) and the integer counters are all-reduced (pairs shard, counters reduce)."""
    from . import parallel
    dev = _device()
    truth = truth or prepare_truth(QUERY_SETS)
    R = truth.R
    rank, world = parallel.world(group)
    mine = list(range(rank, R, world))                 # this rank's database runs (round-robin: equal sizes in practice)
    q = torch.cat([_as_dev(v) for v in QUERY_VECTORS], 0)
    q_run, truth_off, truth_idx = truth.device(dev)
    hist = torch.zeros(R, R, recall_num, device=dev, dtype=torch.int32)
    n_eval = torch.zeros(R, R, device=dev, dtype=torch.int32)
    n_one = torch.zeros(R, R, device=dev, dtype=torch.int32)
    sim = torch.full((q.shape[0], R), float("nan"), device=dev, dtype=torch.float32)
    if mine:
        db = torch.cat([_as_dev(DATABASE_VECTORS[m]) for m in mine], 0)
        sizes = [len(DATABASE_VECTORS[m]) for m in mine]
        seg_off = torch.tensor(np.concatenate(([0], np.cumsum(sizes))), dtype=torch.int32, device=dev)
        seg_run = torch.tensor(mine, dtype=torch.int32, device=dev)
        thresh = torch.tensor([_threshold(n) for n in sizes], dtype=torch.int32, device=dev)
        k = min(recall_num, min(sizes))
        idx, _ = ops.retrieval_tc(db, q, k, seg_off, want_dist=False)                   # [S_local, Nq, k], segment-local rows
        h, ne, no, sm = ops.recall_count(idx, q_run, R, truth_off, truth_idx, thresh, seg_run, db, seg_off, q)
        cols = seg_run.long()
        hist[:, cols] = h
        n_eval[:, cols] = ne
        n_one[:, cols] = no
        sim[:, cols] = sm
    if world > 1:
        packed = torch.cat((hist.reshape(-1), n_eval.reshape(-1), n_one.reshape(-1)))
        parallel.allreduce_counters(packed, group)
        nh, ne_ = hist.numel(), n_eval.numel()
        hist, n_eval, n_one = packed[:nh].view_as(hist), packed[nh: nh + ne_].view_as(n_eval), packed[nh + ne_:].view_as(n_one)
        sim = parallel.merge_nan(sim, group)
    hist_h, ne_h, no_h = hist.cpu().numpy().astype(np.float64), n_eval.cpu().numpy().astype(np.float64), n_one.cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        recall = np.cumsum(hist_h, axis=2) / ne_h[:, :, None] * 100.0
        one_pct = no_h / ne_h * 100.0
    return {"recall": recall, "one_pct": one_pct, "sim": sim.cpu().numpy(), "n_eval": ne_h, "truth": truth}


def get_recall(m, n, DATABASE_VECTORS, QUERY_VECTORS, QUERY_SETS):
    """Same contract as the reference (:162-206): returns (recall[25] cumulative %, top1_similarity_score list,
    one_percent_recall %) for database run m and query run n."""
    dev = _device()
    db, q = _as_dev(DATABASE_VECTORS[m]), _as_dev(QUERY_VECTORS[n])
    Nq, Ndb = q.shape[0], db.shape[0]
    flat, offs = [], [0]
    for i in range(Nq):
        t = QUERY_SETS[n][i][m]
        if len(t):
            flat.extend(t)
        offs.append(len(flat))
    truth_off = torch.tensor(offs, dtype=torch.int32, device=dev)
    truth_idx = torch.tensor(flat if flat else [0], dtype=torch.int32, device=dev)
    seg_off = torch.tensor([0, Ndb], dtype=torch.int32, device=dev)
    k = min(recall_num, Ndb)
    if Nq * Ndb >= ops.RETRIEVAL_TC_MIN and db.shape[1] % 4 == 0:
        idx, _ = ops.retrieval_tc(db, q, k, seg_off, want_dist=False)
    else:
        idx = ops.retrieval_topk(db, q, k, want_dist=False)[0].unsqueeze(0)
    q_run = torch.full((Nq,), -1, dtype=torch.int32, device=dev)           # no own-run skip inside get_recall (the pair loop does it)
    thresh = torch.tensor([_threshold(len(DATABASE_VECTORS[m]))], dtype=torch.int32, device=dev)
    hist, n_eval, n_one, sim = ops.recall_count(idx.contiguous(), q_run, 1, truth_off, truth_idx, thresh, None, db, seg_off, q)
    num_evaluated = float(n_eval[0, 0].item())
    recall = (np.cumsum(hist[0, 0].cpu().numpy()) / num_evaluated) * 100
    s = sim[:, 0].cpu().numpy()
    top1_similarity_score = [v for v in s if not np.isnan(v)]
    one_percent_recall = (float(n_one[0, 0].item()) / num_evaluated) * 100
    return recall, top1_similarity_score, one_percent_recall


class _EmbedGraph:
    """One CUDA graph of `model(x)` for a fixed [batch, 1, N, 3] input (eval mode): the ~25 kernel launches of a step
    replay without any host-side issue cost.  Invalidated when a parameter / buffer of the model changes.  The graph keeps
    the intermediates of one batch alive in its private memory pool (~3 GB for 64 x 4096 points) for as long as the model
    holds it; `get_latent_vectors(..., use_graph=False)` or `del model._lpd_embed_graph` releases / avoids that."""

    def __init__(self, model, shape, dev):
        self.key = self._key(model)
        self.shape = tuple(shape)
        self.x = torch.zeros(shape, device=dev, dtype=torch.float32)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                      # warm-up outside the capture: weight folding, smem opt-ins, allocator
                model(self.x)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.y = model(self.x)

    @staticmethod
    def _key(model):
        # parameters / buffers (torch-side and raw-pointer updates) + everything else the capture bakes in
        return (ops.weights_epoch(), ops.dispatch_key()) + tuple((t.data_ptr(), t._version)
                                                                 for t in list(model.parameters()) + list(model.buffers()))

    def valid_for(self, model, shape):
        return self.shape == tuple(shape) and self.key == self._key(model)


class _Staging:
    """Two-slot staging ring of the embedding driver, cached on the model: pinned host buffers (cudaHostAlloc per batch was
    the dominant and most erratic cost of the end-to-end path) and their device twins."""

    def __init__(self, shape, dev):
        self.shape = tuple(shape)
        self.host = [torch.empty(shape, dtype=torch.float32, pin_memory=True) for _ in range(2)]
        self.dev = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(2)]
        self.copied = [None, None]      # H2D copy of the slot finished (copy stream)
        self.consumed = [None, None]    # compute stream has read the slot's device buffer
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.out_host = None            # pinned landing buffer of the descriptors (grown on demand, copied out on return)

    def out(self, n, dim):
        if self.out_host is None or self.out_host.shape[0] < n or self.out_host.shape[1] != dim:
            self.out_host = torch.empty(max(n, 256), dim, dtype=torch.float32, pin_memory=True)
        return self.out_host[:n]


def get_latent_vectors(model, clouds, batch_num: int = 64, pin: bool = True, use_graph: bool = True) -> np.ndarray:
    """Eval-mode embedding of `clouds` ([n, N, 3] numpy / tensor, host memory) in batches of `batch_num`
    (reference :96-159: batch = eval_batch_size * (1 + P + Nn), tail batch handled, model.eval()/train() toggled).
    Batches go through a two-slot pinned staging ring and a side copy stream, so the host->device copy of batch i+1
    overlaps the kernels of batch i; full batches replay one captured CUDA graph (the ragged tail runs eagerly)."""
    was_training = model.training
    model.eval()
    dev = _device()
    x = clouds if isinstance(clouds, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(clouds, dtype=np.float32))
    x = x.float()
    if x.dim() == 3:
        x = x.unsqueeze(1)
    n = x.shape[0]
    dim = model.net_vlad.output_dim if hasattr(model, "net_vlad") else 256
    if n == 0:
        model.train(was_training)
        return np.empty((0, dim), dtype=np.float32)
    main = torch.cuda.current_stream(dev)
    shape = (batch_num,) + tuple(x.shape[1:])

    st = getattr(model, "_lpd_staging", None)
    if st is None or st.shape != shape or st.dev[0].device != dev:
        st = _Staging(shape, dev)
        model._lpd_staging = st
    copy_stream = st.copy_stream
    out = st.out(n, dim)                                      # cudaHostAlloc per call would cost ~1 ms
    graph = None
    if use_graph and n >= 3 * batch_num:
        graph = getattr(model, "_lpd_embed_graph", None)
        if graph is None or not graph.valid_for(model, shape):
            graph = _EmbedGraph(model, shape, dev)
            model._lpd_embed_graph = graph

    def stage(lo, slot):
        hi = min(n, lo + batch_num)
        cnt = hi - lo
        if st.copied[slot] is not None:
            st.copied[slot].synchronize()                     # the pinned slot is free again (its last H2D copy is done)
        src = x[lo:hi]
        if not (pin and x.is_pinned()):
            st.host[slot][:cnt].copy_(src)                    # pageable -> pinned (host memcpy)
            src = st.host[slot][:cnt]
        with torch.cuda.stream(copy_stream):
            if st.consumed[slot] is not None:
                copy_stream.wait_event(st.consumed[slot])     # the device slot has been read by the compute stream
            st.dev[slot][:cnt].copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        st.copied[slot] = ev
        return lo, hi, slot

    with torch.no_grad():
        staged = stage(0, 0)
        while staged is not None:
            lo, hi, slot = staged
            staged = stage(hi, slot ^ 1) if hi < n else None
            main.wait_event(st.copied[slot])
            d = st.dev[slot][:hi - lo]
            if graph is not None and hi - lo == batch_num:
                graph.x.copy_(d, non_blocking=True)           # device-to-device into the graph's static input
                graph.graph.replay()
                o = graph.y
            else:
                o = model(d)
            done = torch.cuda.Event()
            done.record(main)
            st.consumed[slot] = done
            out[lo:hi].copy_(o, non_blocking=True)
    torch.cuda.synchronize(dev)
    st.copied = [None, None]
    st.consumed = [None, None]
    model.train(was_training)
    return out.numpy().copy()                                 # the landing buffer is reused by the next call


def evaluate_sets(DATABASE_VECTORS, QUERY_VECTORS, QUERY_SETS, truth: Truth | None = None, group=None):
    """Pair loop + aggregation of evaluate_model (reference :59-93) over already-embedded sets, all pairs in one device pass:
    returns (ave_recall[25], average_similarity, ave_one_percent_recall).  ave_recall is the recall@1..25 curve averaged over
    the pairs (the reference's `recall / count`); evaluate_model() reduces it to the reference's scalar."""
    res = recall_all_pairs(DATABASE_VECTORS, QUERY_VECTORS, QUERY_SETS, truth, group)
    R = res["recall"].shape[0]
    off = ~np.eye(R, dtype=bool)                                   # [n][m], m != n
    ave_recall = res["recall"][off].sum(0) / off.sum()
    sims = res["sim"]
    average_similarity = float(np.mean(sims[~np.isnan(sims)].astype(np.float64))) if (~np.isnan(sims)).any() else float("nan")
    ave_one_percent_recall = float(np.mean(res["one_pct"][off]))
    return ave_recall, average_similarity, ave_one_percent_recall


def evaluate_model(model, DATABASE_CLOUDS, QUERY_CLOUDS, QUERY_SETS, batch_num: int = 64, group=None):
    """The reference's evaluate_model (:33-93) over in-memory submaps: DATABASE_CLOUDS[r] / QUERY_CLOUDS[r] are the [n, N, 3]
    clouds of run r (the reference reads them from the pickled file lists, :51-56).  Returns the reference's triple
    (ave_recall SCALAR = mean over N of the mean recall@N, :73; average_similarity_score; ave_one_percent_recall).
    Like the reference (get_latent_vectors :156), the model is left in train() mode."""
    DATABASE_VECTORS = [get_latent_vectors(model, c, batch_num=batch_num) for c in DATABASE_CLOUDS]
    QUERY_VECTORS = [get_latent_vectors(model, c, batch_num=batch_num) for c in QUERY_CLOUDS]
    model.train()
    curve, average_similarity, ave_one_percent_recall = evaluate_sets(DATABASE_VECTORS, QUERY_VECTORS, QUERY_SETS, group=group)
    return float(np.mean(curve)), average_similarity, ave_one_percent_recall
