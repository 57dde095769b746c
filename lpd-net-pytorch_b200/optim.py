"""Fused Adam over flat parameter / gradient buffers + the data-parallel gradient all-reduce.

Replaces `optim.Adam(parameters, learning_rate)` + `optimizer.step()` of the reference
(train_pointnetvlad.py:57,130,159) and the gradient reduction nn.DataParallel does implicitly (:81).

All parameters of a group live in ONE flat fp32 buffer (each nn.Parameter is re-pointed to a view of it), and so do
their gradients, so a step is one NCCL all-reduce (world size > 1) plus ONE lpd_adam launch, instead of ~60 small
kernels per tensor.  The class derives from torch.optim.Optimizer only for the bookkeeping API (param_groups, lr
schedulers such as the reference's ReduceLROnPlateau, state_dict in torch.optim.Adam's own format so checkpoints
written by either implementation load in the other); the update arithmetic is lpd_adam.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from ._lib import LpdError

__all__ = ["Adam"]


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None,
                 allreduce: bool = True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.process_group = process_group
        self.allreduce = allreduce
        self._flat = {}          # group index -> dict(w, g, m, v, views)
        self._step = 0

    # ---- flat buffers -------------------------------------------------------------------------------------------------
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
                views.append((p, off, n, gv))
                off += sz
        flat = dict(w=w, g=g, m=m, v=v, views=views)
        self._flat[gi] = flat
        return flat

    def zero_grad(self, set_to_none: bool = False):
        """Zero the flat gradient buffers (gradients stay views of them; set_to_none is ignored on purpose)."""
        for gi, group in enumerate(self.param_groups):
            flat = self._flat.get(gi) or self._flatten(gi, group)
            if flat is None:
                continue
            flat["g"].zero_()
            for p, _, _, gv in flat["views"]:
                p.grad = gv

    def flat_grads(self):
        """The flat fp32 gradient buffer(s): what the data-parallel all-reduce moves (one NCCL call each)."""
        out = []
        for gi, group in enumerate(self.param_groups):
            flat = self._flat.get(gi) or self._flatten(gi, group)
            if flat is not None:
                out.append(flat["g"])
        return out

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        world = 1
        if self.allreduce and dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.process_group)
        for gi, group in enumerate(self.param_groups):
            flat = self._flat.get(gi) or self._flatten(gi, group)
            if flat is None:
                continue
            steps = flat.setdefault("steps", [0] * len(flat["views"]))
            active = []
            for vi, (p, off, n, gv) in enumerate(flat["views"]):
                if p.grad is None:
                    # torch.optim.Adam skips such a parameter entirely (no moment decay, no weight decay, no step count);
                    # its slice of the flat gradient stays zero for the all-reduce
                    gv.zero_()
                    active.append(False)
                    continue
                if p.grad.data_ptr() != gv.data_ptr():
                    gv.copy_(p.grad)                                   # a caller replaced .grad: bring it into the flat buffer
                    p.grad = gv
                steps[vi] += 1
                active.append(True)
            self._reduce(flat, world)
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

    def _reduce(self, flat, world):
        """the data-parallel gradient sum: one NCCL all-reduce of the flat buffer on the current stream"""
        if world > 1:
            dist.all_reduce(flat["g"], op=dist.ReduceOp.SUM, group=self.process_group)

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
            flat = self._flat.get(gi) or self._flatten(gi, group)
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
