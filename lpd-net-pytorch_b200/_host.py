"""Host-side helpers shared by the module shells: BN folding, weight views, prepared-weight caching.

torch is used here only for device memory (torch.empty / views) and parameter storage; every arithmetic
step goes through lpdnet_b200.ops (the C ABI)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import LpdError


def require_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise LpdError(f"{what}: input is on {x.device}; the lpd_b200 kernels run on sm_100a only "
                       f"(no CPU / torch fallback) — move the module and its input to cuda")
    if x.dtype != torch.float32:
        raise LpdError(f"{what}: expected float32 input, got {x.dtype}")
    if x.device.index != torch.cuda.current_device():
        raise LpdError(f"{what}: input is on {x.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; the "
                       f"kernels launch on the current device's stream — use one process per GPU (torch.cuda.set_device) "
                       f"instead of nn.DataParallel")


def w2d(weight: torch.Tensor) -> torch.Tensor:
    """conv / linear weight [Cout, Cin, 1(,1 or ,3)] -> contiguous [Cout, Cin*...] view (no copy when contiguous)."""
    return weight.detach().reshape(weight.shape[0], -1).contiguous()


def fold_bn(bn: nn.modules.batchnorm._BatchNorm, bias: torch.Tensor | None = None):
    """(scale, shift) of an eval-mode BatchNorm with an optional preceding conv/linear bias."""
    if bn.running_mean is None:
        raise LpdError("BatchNorm without running statistics is not supported")
    return ops.bn_fold(bn.weight.detach() if bn.affine else None, bn.bias.detach() if bn.affine else None,
                       bn.running_mean, bn.running_var, bn.eps, None if bias is None else bias.detach())


class Prepared:
    """Per-module cache of folded / re-laid-out weights, invalidated when any parameter or buffer changes
    (in-place update bumps tensor._version; load_state_dict copies in place; .cuda() replaces tensors; the library's own
    raw-pointer updates — lpd_adam, the running statistics of lpd_bn_finalize — bump ops.weights_epoch())."""

    def __init__(self):
        self._key = None
        self._val = None

    def get(self, module: nn.Module, builder):
        key = (ops.weights_epoch(),) + tuple((t.data_ptr(), t._version)
                                             for t in list(module.parameters()) + list(module.buffers()))
        if key != self._key:
            with torch.no_grad():
                self._val = builder()
            self._key = key
        return self._val


def training_unsupported(module: nn.Module, what: str):
    """Sub-modules called on their own support eval mode only; train mode (batch-statistics BatchNorm + the hand-written
    backward) is entered through PointNetVlad.forward (lpdnet_b200/train.py)."""
    if module.training:
        raise LpdError(
            f"{what}: called directly in train() mode; the training path (batch-statistics BatchNorm + backward) runs "
            f"through PointNetVlad.forward — call .eval() for a stand-alone forward of this sub-module")
