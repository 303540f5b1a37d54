"""Drop-in for the reference's perturbation interface, Classification/attack_algo.py.

Same names, positional order, defaults and return semantics as the reference:
    PGD(x, loss_fn, y=None, model=None, steps=3, gamma=None, start_idx=1, layer_number=16,
        eps=2/255, randinit=False, clip=False) -> x_adv        (attack_algo.py:38-58)
    tensor_clamp / linfball_proj / l2ball_proj                  (attack_algo.py:9-36)
x_adv is returned as a LEAF tensor with requires_grad=True on x's device, like the reference's
`Variable(x_adv, requires_grad=True)`.  Inside, every per-step elementwise span of the reference
(sign, mul, add_, sub, add, lt, gt, nonzero, index, index_put_: 13 launches + 4 host syncs) is ONE
launch of the fused sm_100a kernel; results are bit-identical (tests/test_gpu_pgd.py).

Keyword-only extras (defaults reproduce the reference):
    noise        the U[0,1) draw for the random start.  None -> torch.rand(x.shape) on the CPU generator,
                 exactly where the reference draws it (attack_algo.py:44), then copied to the device.
    rng          'reference' (default) | 'philox': on-device Philox4x32-10 start, no noise tensor in HBM
    seed, offset, offset_device   Philox stream position (rng='philox')
    norm         'linf' (reference) | 'l2' (L2-normalised step + l2ball_proj when clip)
    out_delta    optional tensor that receives x_adv - x after the last step (fused)
    norms_out    optional float32 [2, N] tensor receiving per-sample ||delta||_2, ||delta||_inf after the
                 last step (replaces the D2H + CPU norms of main_perturb.py:188-192)
"""
from typing import Callable, Optional

import torch

from . import ops
from ._lib import AfanError

__all__ = ["PGD", "tensor_clamp", "linfball_proj", "l2ball_proj", "pgd_loop"]


def _as_cuda_f32(x: torch.Tensor, name: str) -> torch.Tensor:
    if not x.is_cuda:
        raise AfanError(f"{name} must live on a CUDA device: afan_b200 has no CPU path")
    return x.detach().contiguous()


def linfball_proj(center, radius, t, in_place=True):
    """attack_algo.py:35-36: clamp t into [center - radius, center + radius] (NaN left alone)."""
    res = t if in_place else t.clone()
    data = res.data if res.is_contiguous() else None
    if data is None:
        raise AfanError("linfball_proj needs a contiguous tensor")
    ops.pgd_linf_step_(None, _as_cuda_f32(center, "center"), data, 0.0, float(radius), True)
    return res


def tensor_clamp(t, min, max, in_place=True):
    """attack_algo.py:9-19 with tensor bounds: t < min -> min, then t > max -> max (NaN left alone), one launch."""
    from . import _lib
    res = t if in_place else t.clone()
    if not res.is_contiguous():
        raise AfanError("tensor_clamp needs a contiguous tensor")
    lo, hi = _as_cuda_f32(min, "min"), _as_cuda_f32(max, "max")
    if lo.shape != res.shape or hi.shape != res.shape:
        raise AfanError("min / max must have the shape of t")
    _lib.check(_lib.lib().afan_tensor_clamp_f32(_lib.f32(res.data, "t"), _lib.f32(lo), _lib.f32(hi), res.numel(),
                                                _lib.stream()), "afan_tensor_clamp_f32")
    return res


def l2ball_proj(center, radius, t, in_place=True):
    """attack_algo.py:21-33 (t == center yields NaN like the reference's 0/0)."""
    res = t if in_place else t.clone()
    if not res.is_contiguous():
        raise AfanError("l2ball_proj needs a contiguous tensor")
    ops.l2ball_proj_(_as_cuda_f32(center, "center"), float(radius), res.data)
    return res


def pgd_loop(x: torch.Tensor, tail_loss: Callable[[torch.Tensor], torch.Tensor], steps: int, gamma: float,
             eps: float, randinit: bool, clip: bool, *, noise=None, rng: str = "reference", seed: Optional[int] = None,
             offset: int = 0, offset_device=None, norm: str = "linf", out_delta=None, norms_out=None,
             workspace=None, x_adv_out=None) -> torch.Tensor:
    """Shared ascent loop of the three reference flavours.  tail_loss(x_adv) -> scalar loss."""
    if norm not in ("linf", "l2"):
        raise AfanError(f"unknown norm {norm!r}")
    x = _as_cuda_f32(x, "x")
    x_adv = torch.empty_like(x) if x_adv_out is None else x_adv_out
    if randinit:
        if noise is None and rng == "reference":
            noise = torch.rand(x.shape)                      # CPU generator, as attack_algo.py:44
        if noise is not None:
            ops.pgd_init(x, eps, noise=noise.to(x.device, non_blocking=True).contiguous(), out=x_adv)
        elif rng == "philox":
            if seed is None:
                raise AfanError("rng='philox' needs seed=")
            ops.pgd_init(x, eps, seed=seed, offset=offset, offset_device=offset_device, out=x_adv)
        else:
            raise AfanError(f"unknown rng {rng!r}")
    else:
        x_adv.copy_(x)
    x_adv.requires_grad_(True)
    need_x = clip or out_delta is not None or norms_out is not None
    for t in range(steps):
        loss = tail_loss(x_adv)
        grad = torch.autograd.grad(loss, x_adv, only_inputs=True)[0]
        last = t == steps - 1
        if norm == "linf":
            ops.pgd_linf_step_(grad.contiguous(), x if need_x else None, x_adv.data, gamma, eps, clip,
                               delta_out=out_delta if last else None, norms_out=norms_out if last else None,
                               workspace=workspace)
        else:
            ops.pgd_l2_step_(grad.contiguous(), x, x_adv.data, gamma, eps, clip, workspace=workspace)
            if last and (out_delta is not None or norms_out is not None):
                ops.pgd_linf_step_(None, x, x_adv.data, 0.0, 0.0, False, delta_out=out_delta, norms_out=norms_out,
                                   workspace=workspace)
    if steps == 0 and (out_delta is not None or norms_out is not None):
        ops.pgd_linf_step_(None, x, x_adv.data, 0.0, 0.0, False, delta_out=out_delta, norms_out=norms_out,
                           workspace=workspace)
    return x_adv


def PGD(x, loss_fn, y=None, model=None, steps=3, gamma=None, start_idx=1, layer_number=16, eps=(2 / 255),
        randinit=False, clip=False, **extras):
    """Classification/attack_algo.py:38-58.  model(x_adv, end_point=layer_number, start_point=start_idx)."""
    def tail_loss(x_adv):
        return loss_fn(model(x_adv, end_point=layer_number, start_point=start_idx), y)

    return pgd_loop(x, tail_loss, steps, gamma, eps, randinit, clip, **extras)
