"""Dual (grouped-statistics) BatchNorm2d with fused residual-add + ReLU, backed by the sm_100a kernels.

What "dual BN" means for A-FAN: the reference has no special module -- the SAME
nn.BatchNorm2d in the tail sees the adversarial batch and then the clean batch in two separate forward
passes (Classification/main_perturb.py:195-196 through resnet_s.py:54,56,89), i.e. separate batch
statistics, shared weight/bias, one running average updated in pass order.  DualBatchNorm2d computes
exactly that for a concatenated [adv; clean] batch in ONE sweep (`groups=2`), and degenerates to a plain
train-mode BatchNorm2d for `groups=1` (the PGD inner-loop passes).  state_dict keys equal nn.BatchNorm2d's.
"""
import contextlib
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import AfanError


_DEFAULT_GROUPS = 1


@contextlib.contextmanager
def statistic_groups(k: int):
    """Inside this context every DualBatchNorm2d called WITHOUT an explicit `groups=` treats its batch as k statistic groups
    along N -- how a library model whose forward calls `self.bn(x)` (torchvision's Bottleneck / ASPP in the DeepLab tail)
    runs k tail passes as ONE batched pass with per-pass statistics."""
    global _DEFAULT_GROUPS
    old, _DEFAULT_GROUPS = _DEFAULT_GROUPS, int(k)
    try:
        yield
    finally:
        _DEFAULT_GROUPS = old


def convert_batchnorm(module: nn.Module, pooled=()) -> nn.Module:
    """Replace every nn.BatchNorm2d below `module` (in place) by a DualBatchNorm2d that SHARES its Parameter / buffer
    objects (optimizer arenas, state_dict keys and checkpoints are unaffected).  Returns `module`.

    pooled: module types whose BatchNorm normalises globally pooled features ([N, C, 1, 1]: N values per channel), e.g.
    torchvision's ASPPPooling.  Those layers get a GroupedLibraryBatchNorm2d instead: nothing streams there, and with a
    handful of values per channel the gradient carries the factor (1 - xhat^2), which any change of summation order moves
    by ~1e-3 -- the library kernel keeps such layers bit-compatible with the reference's."""
    for name, child in list(module.named_children()):
        if isinstance(child, nn.BatchNorm2d) and not isinstance(child, GroupedLibraryBatchNorm2d):
            if not (child.affine and child.track_running_stats) or child.momentum is None:
                raise AfanError("convert_batchnorm needs affine BatchNorm2d layers with running statistics and a momentum")
            if pooled and isinstance(module, tuple(pooled)):
                setattr(module, name, GroupedLibraryBatchNorm2d.from_batchnorm(child))
            else:
                setattr(module, name, DualBatchNorm2d.from_batchnorm(child))
        elif not isinstance(child, DualBatchNorm2d):
            convert_batchnorm(child, pooled)
    return module


class GroupedLibraryBatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d that honours statistic_groups(k): the batch is k passes stacked along N, each normalised by its own
    call of the library kernel in pass order (= what k separate forward passes do to the running statistics)."""

    @classmethod
    def from_batchnorm(cls, bn: nn.BatchNorm2d) -> "GroupedLibraryBatchNorm2d":
        m = cls(bn.num_features, bn.eps, bn.momentum)
        m.weight, m.bias = bn.weight, bn.bias
        m.running_mean, m.running_var, m.num_batches_tracked = bn.running_mean, bn.running_var, bn.num_batches_tracked
        m.train(bn.training)
        return m

    def forward(self, x):
        k = _DEFAULT_GROUPS
        if k == 1 or not self.training:
            return super().forward(x)
        if x.shape[0] % k:
            raise AfanError(f"batch {x.shape[0]} is not divisible into {k} statistic groups")
        return torch.cat([super(GroupedLibraryBatchNorm2d, self).forward(c) for c in x.chunk(k, dim=0)], dim=0)


class _DualBNTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, ws, groups, eps, momentum, relu, replay,
                process_group, mailbox, grad_out=None):
        x = x.contiguous()
        res = residual.contiguous() if residual is not None else None
        y, save_mean, save_invstd = ops.bn_fwd(x, res, weight, bias, running_mean, running_var, ws, groups=groups,
                                               eps=eps, momentum=momentum, relu=relu, replay=replay,
                                               process_group=process_group, mailbox=mailbox)
        ctx.save_for_backward(x, y if relu else None, weight, save_mean, save_invstd)
        ctx.ws, ctx.groups, ctx.relu, ctx.has_res, ctx.pg, ctx.mailbox = ws, groups, relu, residual is not None, process_group, mailbox
        ctx.grad_out = grad_out
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, save_mean, save_invstd = ctx.saved_tensors
        want_res = ctx.has_res and ctx.needs_input_grad[1]
        direct = ctx.grad_out is not None and ctx.needs_input_grad[2] and ctx.needs_input_grad[3]
        dx, dres, dw, db = ops.bn_bwd(dy.contiguous(), x, y, weight, save_mean, save_invstd, ctx.ws,
                                      groups=ctx.groups, relu=ctx.relu, want_dresidual=want_res,
                                      process_group=ctx.pg, mailbox=ctx.mailbox,
                                      dweight_out=ctx.grad_out[0] if direct else None,
                                      dbias_out=ctx.grad_out[1] if direct else None)
        if direct:                       # already stored in weight.grad / bias.grad: nothing for autograd to accumulate
            dw = db = None
        return (dx, dres, dw if ctx.needs_input_grad[2] else None, db if ctx.needs_input_grad[3] else None,
                None, None, None, None, None, None, None, None, None, None, None)


class _AffineEvalFn(torch.autograd.Function):
    """Frozen statistics: y = relu?(x*scale + shift (+res)) from running statistics -- model.eval() of the Classification
    flavour (main_perturb.py:232-246) and the always-frozen BatchNorm of the Detection flavour (Detection/model.py:27-35,
    47-48), where it is differentiated: dx = scale * dy_eff, dres = dy_eff (one fused pass each way)."""

    @staticmethod
    def forward(ctx, x, residual, scale_shift, relu):
        x = x.contiguous()
        res = residual.contiguous() if residual is not None else None
        y = ops.bn_affine(x, res, scale_shift, relu=relu)
        ctx.save_for_backward(y if relu else None, scale_shift)
        ctx.relu = bool(relu)
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        y, scale_shift = ctx.saved_tensors
        dx, dres = ops.bn_affine_bwd(dy.contiguous(), y, scale_shift, relu=ctx.relu,
                                     want_dresidual=ctx.has_res and ctx.needs_input_grad[1])
        return dx, dres, None, None


def frozen_table(bn: nn.Module) -> torch.Tensor:
    """(scale, shift) [C, 2] of a BatchNorm whose statistics AND affine are frozen; cached on the module, rebuilt when any
    of the four tensors is written (load_state_dict) or moved."""
    src = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple(t._version for t in src) + tuple(t.data_ptr() for t in src)
    cached = getattr(bn, "_afan_frozen_table", None)
    if cached is None or cached[0] != key:
        with torch.no_grad():
            scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
            table = torch.stack((scale, bn.bias - bn.running_mean * scale), dim=1).float().contiguous()
        bn._afan_frozen_table = cached = (key, table)
    return cached[1]


def frozen_bn_act(x, bn: nn.Module, residual=None, relu: bool = False):
    """relu?(frozen_bn(x) (+ residual)) in ONE launch (forward) / ONE launch (backward)."""
    return _AffineEvalFn.apply(x, residual, frozen_table(bn), relu)


class DualBatchNorm2d(nn.Module):
    """BatchNorm2d whose forward takes `groups` (statistic groups along the batch), an optional residual
    and a fused ReLU.  Parameters / buffers are named like nn.BatchNorm2d so reference checkpoints load."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self._ws = {}                  # groups -> workspace tensor (device scratch owned by this module)
        self._pending_batches = 0      # host-side count, folded into num_batches_tracked lazily (no launch per pass)
        self.process_group = None      # set by the trainer for NCCL-synchronised statistics
        self.mailbox = None            # p2p.PeerMailbox: statistics exchanged over NVLink inside the kernel instead
        # set by a trainer that owns a zeroed gradient arena AND runs this layer once per differentiated pass: the
        # backward kernel stores d(weight), d(bias) straight into .grad (no temporaries, no accumulation launches)
        self.grad_direct = False
        self._register_state_dict_hook(_flush_hook)

    @classmethod
    def from_batchnorm(cls, bn: nn.BatchNorm2d) -> "DualBatchNorm2d":
        m = cls(bn.num_features, bn.eps, bn.momentum)
        m.weight, m.bias = bn.weight, bn.bias                        # the same Parameter objects
        m.running_mean, m.running_var, m.num_batches_tracked = bn.running_mean, bn.running_var, bn.num_batches_tracked
        m.train(bn.training)
        return m

    def _workspace(self, groups: int, device):
        ws = self._ws.get((groups, device))
        if ws is None:
            ws = self._ws[(groups, device)] = ops.bn_workspace(groups, self.num_features, device)
        return ws

    def flush_batches_tracked(self, multiplier: int = 1):
        if self._pending_batches:
            self.num_batches_tracked += self._pending_batches * multiplier
            self._pending_batches = 0

    def forward(self, x, residual: Optional[torch.Tensor] = None, relu: bool = False, groups: Optional[int] = None,
                replay: int = 1):
        if x.dim() != 4 or x.shape[1] != self.num_features:
            raise AfanError(f"expected [N, {self.num_features}, H, W], got {tuple(x.shape)}")
        if groups is None:
            groups = _DEFAULT_GROUPS
        if self.training:
            self._pending_batches += groups * replay
            grad_out = None
            if self.grad_direct and self.weight.requires_grad and self.bias.requires_grad \
                    and self.weight.grad is not None and self.bias.grad is not None and torch.is_grad_enabled():
                grad_out = (self.weight.grad, self.bias.grad)
            return _DualBNTrainFn.apply(x, residual, self.weight, self.bias, self.running_mean, self.running_var,
                                        self._workspace(groups, x.device), groups, self.eps, self.momentum, relu,
                                        replay, self.process_group, self.mailbox, grad_out)
        invstd = torch.rsqrt(self.running_var + self.eps)
        scale = self.weight.detach() * invstd
        scale_shift = torch.stack((scale, self.bias.detach() - self.running_mean * scale), dim=1).contiguous()
        return _AffineEvalFn.apply(x, residual, scale_shift, relu)

    def extra_repr(self):
        return f"{self.num_features}, eps={self.eps}, momentum={self.momentum}"


def _flush_hook(module, state_dict, prefix, local_metadata):
    module.flush_batches_tracked()
    state_dict[prefix + "num_batches_tracked"] = module.num_batches_tracked.detach().clone()
    return state_dict
