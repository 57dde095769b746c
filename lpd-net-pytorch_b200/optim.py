"""Fused Adam over flat parameter / gradient buffers + the data-parallel gradient all-reduce, bucketed and overlapped with
the backward pass.

Replaces `optim.Adam(parameters, learning_rate)` + `optimizer.step()` of the reference (train_pointnetvlad.py:57,130,159) and
the gradient reduction nn.DataParallel does implicitly (:79-81: replicas' gradients are summed into the master copy).

All parameters of a group live in ONE flat fp32 buffer (each nn.Parameter is re-pointed to a view of it), and so do their
gradients.  With torch.distributed initialised (one process per GPU, each rank owns whole tuples):

  * EARLY bucket(s): every parameter with at least `early_numel` elements (for LPD-Net: net_vlad.hidden1_weights, 16.8 M of the
    17.6 M parameters, whose gradient is the first one the hand-written backward completes) registers a gradient sink; the
    backward (lpdnet_b200/train.py) hands the finished gradient over, it is added into its slice of the flat buffer and
    `ncclAllReduce` of that slice starts at once on NCCL's own stream — it runs under the remaining ~8 ms of backward kernels;
  * LATE bucket: whatever is left (0.8 M parameters, 3.3 MB) is all-reduced in step(), in at most two contiguous pieces;
  * ONE lpd_adam launch then updates everything, with 1/world folded in (mean of the per-rank mean losses).

The class derives from torch.optim.Optimizer only for the bookkeeping API (param_groups, lr schedulers such as the reference's
ReduceLROnPlateau, state_dict in torch.optim.Adam's own format so checkpoints written by either implementation load in the
other); the update arithmetic is lpd_adam.
"""
from __future__ import annotations

import weakref

import torch
import torch.distributed as dist

from . import ops
from ._lib import LpdError

__all__ = ["Adam"]


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None,
                 allreduce: bool = True, overlap: bool = True, early_numel: int = 1 << 20):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.process_group = process_group
        self.allreduce = allreduce
        self.overlap = overlap
        self.early_numel = early_numel
        self.keep_local = False      # diagnostics: keep a copy of this rank's own gradient next to the reduced one
        self._flat = {}              # group index -> dict(w, g, m, v, views, ...)
        self._step = 0

    # ---- flat buffers -------------------------------------------------------------------------------------------------
    def _world(self) -> int:
        if self.allreduce and dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.process_group)
        return 1

    def _flatten(self, gi: int, group: dict):
        ps = [p for p in group["params"] if p.requires_grad]
        if not ps:
            return None
        dev = ps[0].device
        if dev.type != "cuda":
            raise LpdError("lpdnet_b200.optim.Adam: parameters must live on a CUDA device (no CPU fallback)")
        for p in ps:
            if p.dtype != torch.float32 or p.device != dev:
                raise LpdError("lpdnet_b200.optim.Adam: all parameters of a group must be float32 on one device")
        sizes = [((p.numel() + 3) // 4) * 4 for p in ps]             # keep every view 16-byte aligned
        total = sum(sizes)
        w = torch.zeros(total, device=dev, dtype=torch.float32)
        g = torch.zeros(total, device=dev, dtype=torch.float32)
        m = torch.zeros(total, device=dev, dtype=torch.float32)
        v = torch.zeros(total, device=dev, dtype=torch.float32)
        views, off = [], 0
        early = {}
        with torch.no_grad():
            for p, sz in zip(ps, sizes):
                n = p.numel()
                wv = w[off:off + n].view(p.shape)
                wv.copy_(p.data)                                       # one-time move into the flat buffer
                gv = g[off:off + n].view(p.shape)
                p.data = wv
                if p.grad is not None:                                 # a parameter without gradient keeps .grad = None
                    gv.copy_(p.grad)
                    p.grad = gv
                if self.overlap and n >= self.early_numel and self._world() > 1:
                    p._lpd_grad_sink = weakref.ref(self)               # read by lpdnet_b200.train._Grads.add(final=True)
                    early[id(p)] = (gi, len(views))
                views.append((p, off, n, gv))
                off += sz
        flat = dict(w=w, g=g, m=m, v=v, views=views, early=early, pending={}, local=None, reduced=False)
        self._flat[gi] = flat
        return flat

    def _group_of(self, gi):
        return self._flat.get(gi) or self._flatten(gi, self.param_groups[gi])

    def zero_grad(self, set_to_none: bool = False):
        """Zero the flat gradient buffers (gradients stay views of them; set_to_none is ignored on purpose)."""
        for gi in range(len(self.param_groups)):
            flat = self._group_of(gi)
            if flat is None:
                continue
            for work in flat["pending"].values():                      # an early reduction nobody consumed: finish it first
                work.wait()
            flat["pending"].clear()
            flat["reduced"] = False
            flat["g"].zero_()
            for p, _, _, gv in flat["views"]:
                p.grad = gv

    def flat_grads(self):
        """The flat fp32 gradient buffer(s): what the data-parallel all-reduce moves."""
        return [f["g"] for f in (self._group_of(gi) for gi in range(len(self.param_groups))) if f is not None]

    # ---- overlapped reduction -------------------------------------------------------------------------------------------
    def grad_ready(self, param, grad) -> bool:
        """Gradient sink of the hand-written backward: `grad` is the complete gradient of `param` for this backward pass.
        Adds it into the flat buffer and starts the NCCL all-reduce of that slice (async: the collective runs on NCCL's stream
        after the kernels already queued on the current stream, while the backward keeps launching).  Returns True when taken."""
        world = self._world()
        if world == 1:
            return False
        for gi in range(len(self.param_groups)):
            flat = self._flat.get(gi)
            if flat is None or id(param) not in flat["early"]:
                continue
            vi = flat["early"][id(param)][1]
            if vi in flat["pending"]:
                raise LpdError("lpdnet_b200.optim.Adam(overlap=True) reduces a large gradient as soon as the backward finishes it: "
                               "several backward passes per optimizer step (gradient accumulation) need overlap=False")
            p, off, n, gv = flat["views"][vi]
            ops.axpy(gv.view(1, n), n, grad.view(1, n), n, 1, n, 1.0)
            p.grad = gv
            if self.keep_local:
                flat.setdefault("local_early", {})[vi] = gv.detach().clone()
            flat["pending"][vi] = dist.all_reduce(flat["g"][off:off + n], op=dist.ReduceOp.SUM, group=self.process_group, async_op=True)
            return True
        return False

    def _reduce(self, flat, world):
        """finish the data-parallel gradient sum: wait for the early buckets, all-reduce the rest of the flat buffer"""
        if world == 1 or flat["reduced"]:
            return
        if self.keep_local:
            local = flat["g"].detach().clone()
            for vi, t in flat.get("local_early", {}).items():
                _, off, n, _ = flat["views"][vi]
                local[off:off + n].copy_(t.reshape(-1))
            flat["local"] = local
        done = sorted((flat["views"][vi][1], flat["views"][vi][2]) for vi in flat["pending"])
        for work in flat["pending"].values():
            work.wait()                                                # stream-level wait: the current stream continues after NCCL
        flat["pending"].clear()
        flat.pop("local_early", None)
        lo, total = 0, flat["g"].numel()
        for off, n in done + [(total, 0)]:                             # the gaps between the early slices = the late bucket
            if off > lo:
                dist.all_reduce(flat["g"][lo:off], op=dist.ReduceOp.SUM, group=self.process_group)
            lo = max(lo, off + n)
        flat["reduced"] = True

    def reduce_gradients(self):
        """Runs the reduction part of step() now (idempotent until the next zero_grad / step); returns the flat reduced
        gradient of the first group (the SUM over the ranks; lpd_adam applies the 1/world)."""
        world = self._world()
        flats = [self._group_of(gi) for gi in range(len(self.param_groups))]
        for flat in flats:
            if flat is not None:
                self._sync_views(flat)
                self._reduce(flat, world)
        return flats[0]["g"]

    def local_gradients(self):
        """this rank's own flat gradient as it was before the reduction (needs keep_local = True before the backward)"""
        flat = self._flat.get(0)
        if flat is None or flat.get("local") is None:
            raise LpdError("local_gradients(): set optimizer.keep_local = True before the backward pass")
        return flat["local"]

    def describe_reduction(self) -> str:
        world = self._world()
        if world == 1:
            return "single rank, no reduction"
        flat = self._flat.get(0)
        early = sum(flat["views"][vi][2] for _, vi in flat["early"].values()) if flat else 0
        total = sum(v[2] for v in flat["views"]) if flat else 0
        if early:
            return (f"NCCL all-reduce over {world} ranks in 2 buckets: {4 * early / 1e6:.1f} MB (net_vlad.hidden1_weights) launched from inside the "
                    f"backward as soon as it is complete, overlapped with the remaining backward kernels; {4 * (total - early) / 1e6:.1f} MB at step()")
        return f"one NCCL all-reduce of the flat gradient buffer over {world} ranks at step()"

    def _sync_views(self, flat):
        """bring externally assigned .grad tensors into the flat buffer; returns the list of active (has-gradient) flags"""
        steps = flat.setdefault("steps", [0] * len(flat["views"]))
        active = []
        for vi, (p, off, n, gv) in enumerate(flat["views"]):
            if p.grad is None:
                # torch.optim.Adam skips such a parameter entirely (no moment decay, no weight decay, no step count);
                # its slice of the flat gradient stays zero for the all-reduce
                if vi not in flat["pending"]:
                    gv.zero_()
                active.append(vi in flat["pending"])
                continue
            if p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad)                                       # a caller replaced .grad: bring it into the flat buffer
                p.grad = gv
            active.append(True)
        return active, steps

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        world = self._world()
        for gi, group in enumerate(self.param_groups):
            flat = self._group_of(gi)
            if flat is None:
                continue
            active, steps = self._sync_views(flat)
            for vi, a in enumerate(active):
                if a:
                    steps[vi] += 1
            self._reduce(flat, world)
            flat["reduced"] = False
            b1, b2 = group["betas"]
            if all(active) and len(set(steps)) == 1:                   # the common case: ONE launch over the whole buffer
                if steps:
                    ops.adam(flat["w"], flat["g"], flat["m"], flat["v"], group["lr"], b1, b2, group["eps"],
                             group["weight_decay"], steps[0], 1.0 / world)
                continue
            # some parameters received no gradient (or joined later): one launch per run of active segments with equal step
            views = flat["views"]
            vi = 0
            while vi < len(views):
                if not active[vi]:
                    vi += 1
                    continue
                vj = vi
                while vj + 1 < len(views) and active[vj + 1] and steps[vj + 1] == steps[vi]:
                    vj += 1
                lo, hi = views[vi][1], views[vj][1] + views[vj][2]
                ops.adam(flat["w"][lo:hi], flat["g"][lo:hi], flat["m"][lo:hi], flat["v"][lo:hi], group["lr"], b1, b2,
                         group["eps"], group["weight_decay"], steps[vi], 1.0 / world)
                vi = vj + 1
        self._step += 1
        return loss

    # ---- torch.optim.Adam-compatible state ------------------------------------------------------------------------------
    def state_dict(self):
        sd = super().state_dict()
        state, idx = {}, 0
        for gi, group in enumerate(self.param_groups):
            flat = self._flat.get(gi)
            lookup = {} if flat is None else {id(p): (off, n, vi) for vi, (p, off, n, _) in enumerate(flat["views"])}
            steps = [] if flat is None else flat.get("steps", [0] * len(flat["views"]))
            for p in group["params"]:
                if id(p) in lookup and steps[lookup[id(p)][2]] > 0:
                    off, n, vi = lookup[id(p)]
                    state[idx] = {"step": torch.tensor(float(steps[vi])),
                                  "exp_avg": flat["m"][off:off + n].view(p.shape).clone(),
                                  "exp_avg_sq": flat["v"][off:off + n].view(p.shape).clone()}
                idx += 1
        sd["state"] = state
        return sd

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        for group, saved in zip(self.param_groups, groups):
            if "params" in saved and len(saved["params"]) != len(group["params"]):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
        for group, saved in zip(self.param_groups, groups):
            for key in ("lr", "betas", "eps", "weight_decay"):
                if key in saved:
                    group[key] = saved[key]
        idx = 0
        for gi, group in enumerate(self.param_groups):
            flat = self._group_of(gi)
            lookup = {} if flat is None else {id(p): (off, n, vi) for vi, (p, off, n, _) in enumerate(flat["views"])}
            steps = None if flat is None else flat.setdefault("steps", [0] * len(flat["views"]))
            for p in group["params"]:
                st = state_dict["state"].get(idx)
                if st is not None and id(p) in lookup:
                    off, n, vi = lookup[id(p)]
                    flat["m"][off:off + n].copy_(st["exp_avg"].reshape(-1))
                    flat["v"][off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                    steps[vi] = int(float(st["step"]))
                    self._step = max(self._step, steps[vi])
                idx += 1
