"""Tensor-level wrappers over the C ABI (include/afan_b200.h).  CUDA fp32 contiguous tensors only;
every function launches on torch's current stream and never synchronises.  No fallback paths."""
from typing import Optional, Tuple

import torch

from . import _lib, sync
from ._lib import AfanError, check, f32, ptr, stream


def _samples(x: torch.Tensor) -> Tuple[int, int]:
    n = x.shape[0] if x.dim() > 0 else 1
    return n, (x.numel() // n if n > 0 else 0)


# ------------------------------------------------------------------------------------------------
# PGD (a2, a3, a4, a11)
# ------------------------------------------------------------------------------------------------
def pgd_init(x: torch.Tensor, eps: float, *, noise: Optional[torch.Tensor] = None, seed: Optional[int] = None,
             offset: int = 0, offset_device: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Random start x + (2u-1)*eps (Classification/attack_algo.py:42-44).  `noise` = u drawn by the
    caller (bitwise parity with the reference's CPU torch.rand); otherwise on-device Philox(seed, offset)."""
    out = torch.empty_like(x) if out is None else out
    if x.dtype == torch.bfloat16:                            # bf16-storage twin (config 3); noise stays fp32
        bp = lambda t, nm: _lib.dev_ptr(t, torch.bfloat16, nm)
        if noise is not None:
            if noise.shape != x.shape:
                raise AfanError("noise must have the shape of x")
            check(_lib.lib().afan_pgd_init_noise_bf16(bp(x, "x"), f32(noise.float().contiguous(), "noise"), bp(out, "out"),
                                                      x.numel(), float(eps), stream()), "afan_pgd_init_noise_bf16")
        else:
            if seed is None:
                raise AfanError("pgd_init needs either noise= or seed=")
            check(_lib.lib().afan_pgd_init_philox_bf16(bp(x, "x"), bp(out, "out"), x.numel(), float(eps),
                                                       int(seed) & (2 ** 64 - 1), int(offset), ptr(offset_device),
                                                       stream()), "afan_pgd_init_philox_bf16")
        return out
    if noise is not None:
        if noise.shape != x.shape:
            raise AfanError("noise must have the shape of x")
        check(_lib.lib().afan_pgd_init_noise_f32(f32(x, "x"), f32(noise, "noise"), f32(out, "out"), x.numel(),
                                                 float(eps), stream()), "afan_pgd_init_noise_f32")
    else:
        if seed is None:
            raise AfanError("pgd_init needs either noise= or seed=")
        if offset_device is not None and offset_device.dtype not in (torch.int64, torch.uint64):
            raise AfanError("offset_device must be a 64-bit integer CUDA tensor")
        check(_lib.lib().afan_pgd_init_philox_f32(f32(x, "x"), f32(out, "out"), x.numel(), float(eps),
                                                  int(seed) & (2 ** 64 - 1), int(offset), ptr(offset_device),
                                                  stream()), "afan_pgd_init_philox_f32")
    return out


def norms_workspace(n_samples: int, device) -> torch.Tensor:
    nbytes = _lib.lib().afan_pgd_norms_workspace_bytes(int(n_samples))
    return torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device=device)


def pgd_linf_step_(grad: torch.Tensor, x_clean: Optional[torch.Tensor], x_adv: torch.Tensor, gamma: float,
                   eps: float, clip: bool, *, delta_out: Optional[torch.Tensor] = None,
                   norms_out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One fused L-inf PGD update IN PLACE on x_adv (Classification/attack_algo.py:53-56).
    norms_out: float32 [2, N] -> per-sample L2 and Linf of (x_adv - x_clean) (main_perturb.py:188-192)."""
    n, per = _samples(x_adv)
    if (grad is not None and grad.shape != x_adv.shape) or (x_clean is not None and x_clean.shape != x_adv.shape):
        raise AfanError("grad / x_clean must have the shape of x_adv")
    if norms_out is not None:
        if workspace is None:
            workspace = norms_workspace(n, x_adv.device)
        if norms_out.numel() != 2 * n:
            raise AfanError("norms_out must hold 2*N floats")
    if x_adv.dtype == torch.bfloat16:
        bp = lambda t, nm: _lib.dev_ptr(t, torch.bfloat16, nm)
        check(_lib.lib().afan_pgd_linf_step_bf16(
            bp(grad, "grad"), bp(x_clean, "x_clean"), bp(x_adv, "x_adv"), bp(delta_out, "delta_out"),
            f32(norms_out, "norms_out"), ptr(workspace), workspace.numel() * workspace.element_size() if workspace is not None else 0,
            n, per, float(gamma), float(eps), int(bool(clip)), stream()), "afan_pgd_linf_step_bf16")
        return x_adv
    check(_lib.lib().afan_pgd_linf_step_f32(
        f32(grad, "grad"), f32(x_clean, "x_clean"), f32(x_adv, "x_adv"), f32(delta_out, "delta_out"),
        f32(norms_out, "norms_out"), ptr(workspace), workspace.numel() * workspace.element_size() if workspace is not None else 0,
        n, per, float(gamma), float(eps), int(bool(clip)), stream()), "afan_pgd_linf_step_f32")
    return x_adv


def sample_l2norm(a: torch.Tensor, b: Optional[torch.Tensor] = None, *, out: Optional[torch.Tensor] = None,
                  workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-sample ||a - b||_2 (b=None -> ||a||_2): `.view(N,-1).norm(p=2, dim=1)` of attack_algo.py:28."""
    n, per = _samples(a)
    out = torch.empty(n, dtype=torch.float32, device=a.device) if out is None else out
    workspace = norms_workspace(n, a.device) if workspace is None else workspace
    check(_lib.lib().afan_sample_l2norm_f32(f32(a, "a"), f32(b, "b"), f32(out, "out"), ptr(workspace),
                                            workspace.numel() * workspace.element_size(), n, per, stream()),
          "afan_sample_l2norm_f32")
    return out


def pgd_l2_step_(grad, x_clean, x_adv, gamma, eps, clip, *, delta_out=None, tiny=1e-12, workspace=None):
    """L2-normalised ascent IN PLACE (+ l2ball_proj when clip), see include/afan_b200.h a5/a5b."""
    n, per = _samples(x_adv)
    workspace = norms_workspace(n, x_adv.device) if workspace is None else workspace
    L = _lib.lib()
    gnorm = sample_l2norm(grad, workspace=workspace)
    check(L.afan_pgd_l2_step_f32(f32(grad), f32(gnorm), f32(x_adv), n, per, float(gamma), float(tiny), stream()),
          "afan_pgd_l2_step_f32")
    if clip:
        l2ball_proj_(x_clean, eps, x_adv, delta_out=delta_out, workspace=workspace)
    elif delta_out is not None:
        pgd_linf_step_(None, x_clean, x_adv, 0.0, 0.0, False, delta_out=delta_out)   # delta = x_adv - x only
    return x_adv


def l2ball_proj_(center, radius, t, *, delta_out=None, workspace=None):
    """Classification/attack_algo.py:21-33 IN PLACE on t."""
    n, per = _samples(t)
    dist = sample_l2norm(t, center, workspace=workspace)
    check(_lib.lib().afan_l2ball_proj_f32(f32(center), f32(dist), f32(t), f32(delta_out), n, per, float(radius),
                                          stream()), "afan_l2ball_proj_f32")
    return t


# ------------------------------------------------------------------------------------------------
# mix_feature (a8)
# ------------------------------------------------------------------------------------------------
def mix_feature(clean: torch.Tensor, adv: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if clean.shape != adv.shape or clean.dim() < 2:
        raise AfanError("mix_feature needs two [N, C, ...] tensors of the same shape")
    out = torch.empty_like(clean) if out is None else out
    n, c = clean.shape[0], clean.shape[1]
    hw = clean.numel() // max(n * c, 1)
    check(_lib.lib().afan_mix_feature_f32(f32(clean, "clean"), f32(adv, "adv"), f32(out, "out"), n, c, hw, stream()),
          "afan_mix_feature_f32")
    return out


def sat_mix(clean: torch.Tensor, adv: torch.Tensor, weights, mix_flags):
    """Fused SAT points: out_j = mix_flags[j] ? mix_feature(clean, lerp(clean, adv, w_j)) : lerp(clean, adv, w_j)
    for up to 4 points in ONE launch (get_sample_points + per-point mix_feature of the reference)."""
    import ctypes
    m = len(weights)
    if m != len(mix_flags) or not 1 <= m <= 4:
        raise AfanError("sat_mix takes 1..4 points with one mix flag each")
    if clean.shape != adv.shape or clean.dim() < 2:
        raise AfanError("sat_mix needs two [N, C, ...] tensors of the same shape")
    outs = [torch.empty_like(clean) for _ in range(m)]
    n, c = clean.shape[0], clean.shape[1]
    hw = clean.numel() // max(n * c, 1)
    optr = (ctypes.c_void_p * m)(*[f32(o) for o in outs])
    wts = (ctypes.c_float * m)(*[float(w) for w in weights])
    flg = (ctypes.c_int * m)(*[int(bool(f)) for f in mix_flags])
    check(_lib.lib().afan_sat_mix_f32(f32(clean, "clean"), f32(adv, "adv"), optr, wts, flg, m, n, c, hw, stream()),
          "afan_sat_mix_f32")
    return outs


# ------------------------------------------------------------------------------------------------
# dual BatchNorm (a10)
# ------------------------------------------------------------------------------------------------
def bn_workspace(groups: int, channels: int, device) -> torch.Tensor:
    nbytes = _lib.lib().afan_bn_workspace_bytes(int(groups), int(channels))
    if nbytes < 0:
        check(int(nbytes), "afan_bn_workspace_bytes")
    return torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device=device)


def _nchw(x: torch.Tensor, groups: int):
    if x.dim() < 2:
        raise AfanError("BatchNorm input must be [N, C, ...]")
    gn, c = x.shape[0], x.shape[1]
    if gn % groups:
        raise AfanError(f"batch {gn} is not divisible into {groups} statistic groups")
    return gn // groups, c, x.numel() // max(gn * c, 1)


def _ws_bytes(ws):
    return ws.numel() * ws.element_size()


def bn_fwd(x, residual, weight, bias, running_mean, running_var, ws, *, groups=1, eps=1e-5, momentum=0.1,
           relu=False, replay=1, process_group=None, split=False, mailbox=None):
    """Train-mode grouped-statistics BN (+residual, +ReLU).  Returns (y, save_mean[G,C], save_invstd[G,C]).
    With a process group of size > 1 the per-(group, channel) sums are all-reduced over NCCL in ONE
    message for all groups between the statistics kernel and the apply kernel."""
    n, c, hw = _nchw(x, groups)
    L = _lib.lib()
    y = torch.empty_like(x)
    save_mean = torch.empty((groups, c), dtype=torch.float32, device=x.device)
    save_invstd = torch.empty_like(save_mean)
    world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
    if mailbox is not None and not split:                    # fused: statistics exchanged over NVLink inside the kernel
        rc = L.afan_bn_fwd_p2p_f32(f32(x), f32(residual), f32(weight), f32(bias), f32(running_mean), f32(running_var),
                                   f32(y), f32(save_mean), f32(save_invstd), groups, n, c, hw, float(eps), float(momentum),
                                   int(bool(relu)), int(replay), mailbox.world, mailbox.rank, mailbox.peer_ptrs,
                                   mailbox.cmax, ptr(mailbox.state), stream())
        if rc != _lib.AFAN_ERR_UNSUPPORTED:
            check(rc, "afan_bn_fwd_p2p_f32")
            return y, save_mean, save_invstd
        world = mailbox.world                                # shape not cluster-resident: same decision on every rank
    if world == 1 and not split:
        check(L.afan_bn_fwd_f32(f32(x), f32(residual), f32(weight), f32(bias), f32(running_mean), f32(running_var),
                                f32(y), f32(save_mean), f32(save_invstd), ptr(ws), _ws_bytes(ws), groups, n, c, hw,
                                float(eps), float(momentum), int(bool(relu)), int(replay), stream()), "afan_bn_fwd_f32")
    else:
        sums = torch.empty((groups, c, 2), dtype=torch.float64, device=x.device)
        check(L.afan_bn_fwd_stats_f32(f32(x), ptr(sums), ptr(ws), _ws_bytes(ws), groups, n, c, hw, stream()),
              "afan_bn_fwd_stats_f32")
        sync.allreduce_sums_(sums, process_group)             # ONE message: clean + adversarial statistics
        check(L.afan_bn_fwd_finalize_f32(ptr(sums), sync.global_count(n, hw, world), f32(weight), f32(bias), f32(running_mean),
                                         f32(running_var), f32(save_mean), f32(save_invstd), ptr(ws), _ws_bytes(ws),
                                         groups, c, float(eps), float(momentum), int(replay), stream()),
              "afan_bn_fwd_finalize_f32")
        check(L.afan_bn_fwd_apply_f32(f32(x), f32(residual), f32(y), ptr(ws), _ws_bytes(ws), groups, n, c, hw,
                                      int(bool(relu)), stream()), "afan_bn_fwd_apply_f32")
    return y, save_mean, save_invstd


def bn_bwd(dy, x, y, weight, save_mean, save_invstd, ws, *, groups=1, relu=False, want_dresidual=False,
           process_group=None, split=False, mailbox=None, dweight_out=None, dbias_out=None):
    """Backward of bn_fwd.  Returns (dx, dresidual|None, dweight[C], dbias[C]) (dweight/dbias are LOCAL sums;
    the data-parallel gradient all-reduce averages them with the other parameters)."""
    n, c, hw = _nchw(x, groups)
    L = _lib.lib()
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dresidual else None
    # dweight_out / dbias_out: the kernels STORE the sums there (e.g. straight into the gradient arena)
    dweight = torch.empty(c, dtype=torch.float32, device=x.device) if dweight_out is None else dweight_out
    dbias = torch.empty_like(dweight) if dbias_out is None else dbias_out
    world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
    if mailbox is not None and not split:
        rc = L.afan_bn_bwd_p2p_f32(f32(dy), f32(x), f32(y) if relu else None, f32(weight), f32(save_mean), f32(save_invstd),
                                   f32(dx), f32(dres), f32(dweight), f32(dbias), groups, n, c, hw, int(bool(relu)),
                                   mailbox.world, mailbox.rank, mailbox.peer_ptrs, mailbox.cmax, ptr(mailbox.state), stream())
        if rc != _lib.AFAN_ERR_UNSUPPORTED:
            check(rc, "afan_bn_bwd_p2p_f32")
            return dx, dres, dweight, dbias
        world = mailbox.world
    if world == 1 and not split:
        check(L.afan_bn_bwd_f32(f32(dy), f32(x), f32(y) if relu else None, f32(weight), f32(save_mean),
                                f32(save_invstd), f32(dx), f32(dres), f32(dweight), f32(dbias), ptr(ws), _ws_bytes(ws),
                                groups, n, c, hw, int(bool(relu)), stream()), "afan_bn_bwd_f32")
    else:
        sums = torch.empty((groups, c, 2), dtype=torch.float64, device=x.device)
        check(L.afan_bn_bwd_reduce_f32(f32(dy), f32(x), f32(y) if relu else None, f32(save_mean), f32(save_invstd),
                                       ptr(sums), f32(dweight), f32(dbias), ptr(ws), _ws_bytes(ws), groups, n, c, hw,
                                       int(bool(relu)), stream()), "afan_bn_bwd_reduce_f32")
        sync.allreduce_sums_(sums, process_group)
        check(L.afan_bn_bwd_finalize_f32(ptr(sums), sync.global_count(n, hw, world), f32(weight), f32(save_mean),
                                         f32(save_invstd), ptr(ws), _ws_bytes(ws), groups, c, stream()),
              "afan_bn_bwd_finalize_f32")
        check(L.afan_bn_bwd_apply_f32(f32(dy), f32(x), f32(y) if relu else None, f32(dx), f32(dres), ptr(ws),
                                      _ws_bytes(ws), groups, n, c, hw, int(bool(relu)), stream()),
              "afan_bn_bwd_apply_f32")
    return dx, dres, dweight, dbias


def bn_affine(x, residual, scale_shift, *, relu=False):
    """Inference-mode BN: y = relu?(x*scale[c] + shift[c] (+residual)); scale_shift float32 [C, 2]."""
    n, c, hw = _nchw(x, 1)
    y = torch.empty_like(x)
    check(_lib.lib().afan_bn_affine_f32(f32(x), f32(residual), f32(scale_shift), f32(y), n, c, hw, int(bool(relu)),
                                        stream()), "afan_bn_affine_f32")
    return y


def bn_affine_bwd(dy, y, scale_shift, *, relu=False, want_dresidual=False):
    """Backward of bn_affine for frozen statistics: (dx = scale[c] * dy_eff, dresidual = dy_eff or None)."""
    n, c, hw = _nchw(dy, 1)
    dx = torch.empty_like(dy)
    dres = torch.empty_like(dy) if want_dresidual else None
    check(_lib.lib().afan_bn_affine_bwd_f32(f32(dy), f32(y) if relu else None, f32(scale_shift), f32(dx), f32(dres), n, c, hw,
                                            int(bool(relu)), stream()), "afan_bn_affine_bwd_f32")
    return dx, dres


# ------------------------------------------------------------------------------------------------
# fused SGD (a7 tail)
# ------------------------------------------------------------------------------------------------
def sgd_momentum_(param, grad, buf, lr_device, *, momentum=0.9, weight_decay=5e-4, grad_scale=1.0):
    if not (param.numel() == grad.numel() == buf.numel()):
        raise AfanError("param / grad / momentum buffer sizes differ")
    check(_lib.lib().afan_sgd_momentum_f32(f32(param), f32(grad), f32(buf), param.numel(), f32(lr_device),
                                           float(momentum), float(weight_decay), float(grad_scale), stream()),
          "afan_sgd_momentum_f32")
    return param


# ------------------------------------------------------------------------------------------------
# NMS (f4)
# ------------------------------------------------------------------------------------------------
def nms_flags(boxes: torch.Tensor, scores: torch.Tensor, threshold: float):
    """Greedy NMS on the device.  Returns (keep_flags uint8 [N] by original index, count int32 [1]); no host sync."""
    if boxes.dim() != 2 or boxes.shape[1] != 4 or scores.shape[0] != boxes.shape[0]:
        raise AfanError("nms needs boxes [N, 4] and scores [N]")
    n = boxes.shape[0]
    keep = torch.zeros(n, dtype=torch.uint8, device=boxes.device)
    count = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    if n == 0:
        return keep, count
    order = torch.sort(scores, descending=True, stable=True)[1]
    sorted_boxes = boxes.index_select(0, order).contiguous()
    nbytes = _lib.lib().afan_nms_workspace_bytes(n)
    ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=boxes.device)
    check(_lib.lib().afan_nms_f32(f32(sorted_boxes, "boxes"), _lib.dev_ptr(order, torch.int64, "order"), float(threshold),
                                  _lib.dev_ptr(keep, torch.uint8, "keep"), _lib.dev_ptr(count, torch.int32, "count"),
                                  ptr(ws), ws.numel() * 8, n, stream()), "afan_nms_f32")
    return keep, count


class _ShortcutAFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pad):
        x = x.contiguous()
        n, c, h, w = x.shape
        y = torch.empty(n, c + 2 * pad, (h + 1) // 2, (w + 1) // 2, dtype=x.dtype, device=x.device)
        check(_lib.lib().afan_shortcut_a_fwd_f32(f32(x, "x"), f32(y), n, c, h, w, pad, stream()), "afan_shortcut_a_fwd_f32")
        ctx.cfg = (n, c, h, w, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, pad = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty(n, c, h, w, dtype=dy.dtype, device=dy.device)
        check(_lib.lib().afan_shortcut_a_bwd_f32(f32(dy, "dy"), f32(dx), n, c, h, w, pad, stream()), "afan_shortcut_a_bwd_f32")
        return dx, None


def shortcut_a(x: torch.Tensor, pad: int) -> torch.Tensor:
    """Option-A shortcut of a stage transition: F.pad(x[:, :, ::2, ::2], (0, 0, 0, 0, pad, pad)) in one launch each way
    (Classification/resnet_s.py:60-63)."""
    return _ShortcutAFn.apply(x, int(pad))


class _LinearFn(torch.autograd.Function):
    """F.linear whose weight / bias gradient is ONE deterministic launch (afan_linear_wgrad_f32); output and input gradient
    stay library GEMMs (5-6 us each at the classifier's 256 x 64 x 100)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dx = dy.mm(weight) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dy_c, x_c = dy.contiguous(), x.contiguous()
            dw = torch.empty_like(weight)
            db = torch.empty(weight.shape[0], dtype=weight.dtype, device=weight.device) if ctx.has_bias else None
            check(_lib.lib().afan_linear_wgrad_f32(f32(dy_c, "dy"), f32(x_c, "x"), f32(dw), f32(db), x_c.shape[0], weight.shape[0],
                                                   weight.shape[1], stream()), "afan_linear_wgrad_f32")
        return dx, dw, db


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _LinearFn.apply(x, weight, bias)


def nms_batched(boxes_sorted: torch.Tensor, threshold: float, max_keep: int, want_flags: bool = False):
    """Greedy NMS of `images` ranked box lists in one launch pair.  boxes_sorted [B, N, 4] (descending score per image).
    Returns (kept boxes [B, max_keep, 4] in rank order, zero-padded; counts int32 [B]; keep flags uint8 [B, N] by rank or
    None).  No host sync."""
    if boxes_sorted.dim() != 3 or boxes_sorted.shape[2] != 4:
        raise AfanError("nms_batched needs boxes [B, N, 4]")
    b, n = boxes_sorted.shape[:2]
    dev = boxes_sorted.device
    kept = torch.empty(b, max_keep, 4, dtype=torch.float32, device=dev)
    counts = torch.empty(b, dtype=torch.int32, device=dev)
    flags = torch.empty(b, n, dtype=torch.uint8, device=dev) if want_flags else None
    nbytes = _lib.lib().afan_nms_batched_workspace_bytes(b, n)
    ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=dev)
    check(_lib.lib().afan_nms_batched_f32(f32(boxes_sorted, "boxes"), float(threshold), int(max_keep), f32(kept),
                                          _lib.dev_ptr(flags, torch.uint8, "keep") if want_flags else None,
                                          _lib.dev_ptr(counts, torch.int32, "counts"), ptr(ws), ws.numel() * 8, b, n, stream()),
          "afan_nms_batched_f32")
    return kept, counts, flags


# ------------------------------------------------------------------------------------------------
# 3x3 / stride 1 / pad 1 convolutions of the re-executed tail (Classification/resnet_s.py:53,55)
# ------------------------------------------------------------------------------------------------
CONV3X3_CHANNELS = (16, 32, 64)
CONV3X3_SIZES = (8, 16, 32)


def conv3x3_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    """Shapes the hand-written kernels cover: square 8/16/32 maps, C_in == C_out in {16, 32, 64}, fp32 NCHW."""
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and weight.shape[0] == weight.shape[1] == x.shape[1]
            and x.shape[1] in CONV3X3_CHANNELS and x.shape[2] == x.shape[3] and x.shape[2] in CONV3X3_SIZES
            and tuple(weight.shape[2:]) == (3, 3))


CONV_TC_AVAILABLE = True          # the tcgen05 (UMMA) convolution is built into libafan_b200.so


def conv3x3_umma_supported(n: int, c: int, h: int) -> bool:
    """Shapes of the tcgen05 implicit-GEMM convolution: (C, H) in {(32, 16), (64, 8)} (even N for 8x8 maps)."""
    return bool(_lib.lib().afan_conv3x3_umma_supported(int(n), int(c), int(h)))


def conv3x3_pack(descs: torch.Tensor, c_max: int, math: str = "fp32"):
    """Repack every layer listed in `descs` (int64 [L, 4] = weight ptr, fwd-packed ptr, dgrad-packed ptr, C) in ONE launch.
    math: "fp32" (FFMA kernels, C*9*C floats per packing), "tf32" (tensor cores, 1 pass, C*9*C floats) or
    "3xtf32" (tensor cores, hi/lo split, 2*C*9*C floats), or "umma" (tcgen05 3xTF32 implicit GEMM for C in {32, 64},
    2*C*9*C floats; the other layers get the "fp32" packing)."""
    d = _lib.dev_ptr(descs, torch.int64, "descs")
    if math == "fp32":
        check(_lib.lib().afan_conv3x3_pack_f32(d, descs.shape[0], int(c_max), stream()), "afan_conv3x3_pack_f32")
    elif math == "umma":
        # layers the tcgen05 kernel does not cover (C = 16, the stride-2 transitions) keep the FFMA packing
        check(_lib.lib().afan_conv3x3_pack_f32(d, descs.shape[0], int(c_max), stream()), "afan_conv3x3_pack_f32")
        check(_lib.lib().afan_conv3x3_pack_umma_f32(d, descs.shape[0], int(c_max), stream()), "afan_conv3x3_pack_umma_f32")
    else:
        check(_lib.lib().afan_conv3x3_pack_tc_f32(d, descs.shape[0], int(c_max), _TC_PASSES[math], stream()),
              "afan_conv3x3_pack_tc_f32")


_TC_PASSES = {"tf32": 1, "3xtf32": 3}


def conv3x3(x: torch.Tensor, w_packed: torch.Tensor, variant: int = 0, math: str = "fp32",
            addend: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = conv2d(x, W, stride 1, pad 1) (+ addend) with W packed by conv3x3_pack (forward packing), or the input gradient
    of that convolution when given dy and the dgrad packing.  `math` must match the packing."""
    n, c, h, _ = x.shape
    y = torch.empty_like(x)
    if addend is not None and addend.shape != x.shape:
        raise AfanError("addend must have the shape of the output")
    if math == "fp32":
        check(_lib.lib().afan_conv3x3_f32(f32(x, "x"), f32(w_packed, "w_packed"), f32(y), f32(addend, "addend"), n, c, h,
                                          int(variant), stream()), "afan_conv3x3_f32")
    elif math == "umma":
        check(_lib.lib().afan_conv3x3_umma_f32(f32(x, "x"), f32(w_packed, "w_packed"), f32(y), f32(addend, "addend"), n, c, h,
                                               stream()), "afan_conv3x3_umma_f32")
    else:
        check(_lib.lib().afan_conv3x3_tc_f32(f32(x, "x"), f32(w_packed, "w_packed"), f32(y), f32(addend, "addend"), n, c, h,
                                             _TC_PASSES[math], int(variant), stream()), "afan_conv3x3_tc_f32")
    return y


def conv3x3_umma_bn_workspace(n: int, c: int, device) -> torch.Tensor:
    """Per-CTA partial statistics of conv3x3_umma_bn(stats_out=...) (need not be zeroed; one per producer call in flight)."""
    return torch.empty(_lib.lib().afan_conv3x3_umma_bn_workspace_bytes(int(n), int(c)) // 8, dtype=torch.float64, device=device)


def conv3x3_umma_bn(x: torch.Tensor, w_packed: torch.Tensor, *, stats_out: Optional[torch.Tensor] = None,
                    stats_in: Optional[torch.Tensor] = None, bn=None, groups: int = 1, eps: float = 1e-5,
                    momentum: float = 0.1, replay: int = 1, mailbox=None):
    """tcgen05 convolution with BatchNorm folded in (Classification/resnet_s.py:70-72).

    stats_out (producer): workspace that receives the per-CTA {sum, sum of squares} of the OUTPUT.
    stats_in + bn = (weight, bias, running_mean, running_var) (consumer): x is the producer's raw output; its train-mode
    BatchNorm (+ ReLU) over `groups` statistic groups is applied while loading.  Returns (y, save_mean, save_invstd, table)
    -- the last three only for a consumer call (what bn_bwd_xmask needs).
    mailbox (p2p.PeerMailbox, consumer only): statistics of the GLOBAL batch, exchanged with the peer GPUs inside the kernel."""
    n, c, h, _ = x.shape
    y = torch.empty_like(x)
    sm = si = tab = None
    a = [None] * 4
    if stats_in is not None:
        if bn is None:
            raise AfanError("stats_in needs bn=(weight, bias, running_mean, running_var)")
        sm = torch.empty((groups, c), dtype=torch.float32, device=x.device)
        si = torch.empty_like(sm)
        tab = torch.empty((groups, c, 2), dtype=torch.float32, device=x.device)
        a = [f32(t) for t in bn]
    need = _lib.lib().afan_conv3x3_umma_bn_workspace_bytes(n, c) // 8
    for ws in (stats_in, stats_out):
        if ws is not None and (ws.dtype != torch.float64 or ws.numel() < need or not ws.is_cuda):
            raise AfanError("statistics workspace must be a CUDA float64 tensor from conv3x3_umma_bn_workspace(n, c)")
    if mailbox is not None and stats_in is not None:
        if stats_out is not None:
            raise AfanError("the multi-GPU consumer call does not also produce statistics")
        check(_lib.lib().afan_conv3x3_umma_bn_p2p_f32(f32(x, "x"), f32(w_packed, "w_packed"), f32(y), ptr(stats_in), a[0], a[1], a[2],
                                                      a[3], f32(sm), f32(si), f32(tab), int(groups), n, c, h, float(eps),
                                                      float(momentum), int(replay), mailbox.world, mailbox.rank, mailbox.peer_ptrs,
                                                      mailbox.cmax, ptr(mailbox.state), stream()), "afan_conv3x3_umma_bn_p2p_f32")
        return y, sm, si, tab
    check(_lib.lib().afan_conv3x3_umma_bn_f32(f32(x, "x"), f32(w_packed, "w_packed"), f32(y), ptr(stats_in),
                                              a[0], a[1], a[2], a[3], f32(sm), f32(si), f32(tab), ptr(stats_out), int(groups),
                                              n, c, h, float(eps), float(momentum), int(replay), stream()),
          "afan_conv3x3_umma_bn_f32")
    return y, sm, si, tab


def bn_bwd_xmask(dy, x, table, weight, save_mean, save_invstd, *, groups=1, dweight_out=None, dbias_out=None, mailbox=None):
    """Backward of BatchNorm + ReLU whose output was never stored (consumed by conv3x3_umma_bn(in_table=table)): the
    ReLU mask is recomputed from x and the table.  Returns (dx, dweight, dbias)."""
    n, c, hw = _nchw(x, groups)
    dx = torch.empty_like(x)
    dweight = torch.empty(c, dtype=torch.float32, device=x.device) if dweight_out is None else dweight_out
    dbias = torch.empty_like(dweight) if dbias_out is None else dbias_out
    if mailbox is not None:
        check(_lib.lib().afan_bn_bwd_xmask_p2p_f32(f32(dy, "dy"), f32(x, "x"), f32(table, "table"), f32(weight), f32(save_mean),
                                                   f32(save_invstd), f32(dx), f32(dweight), f32(dbias), groups, n, c, hw,
                                                   mailbox.world, mailbox.rank, mailbox.peer_ptrs, mailbox.cmax,
                                                   ptr(mailbox.state), stream()), "afan_bn_bwd_xmask_p2p_f32")
        return dx, dweight, dbias
    check(_lib.lib().afan_bn_bwd_xmask_f32(f32(dy, "dy"), f32(x, "x"), f32(table, "table"), f32(weight), f32(save_mean),
                                           f32(save_invstd), f32(dx), f32(dweight), f32(dbias), groups, n, c, hw, stream()),
          "afan_bn_bwd_xmask_f32")
    return dx, dweight, dbias


CONV3X3S2_SHAPES = ((16, 32), (32, 16))          # (C_in, H_in) of the two stride-2 stage transitions of the CIFAR ResNets


def conv3x3s2_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[2] == x.shape[3]
            and (x.shape[1], x.shape[2]) in CONV3X3S2_SHAPES and tuple(weight.shape) == (2 * x.shape[1], x.shape[1], 3, 3))


def conv3x3s2(x: torch.Tensor, w_packed: torch.Tensor, dgrad: bool = False) -> torch.Tensor:
    """Stride-2 transition convolution (C -> 2C, H -> H/2), or with dgrad=True its input gradient (x is then dy
    [N, 2C, H/2, H/2] and the result [N, C, H, H])."""
    n = x.shape[0]
    if dgrad:
        cin, ho = x.shape[1] // 2, x.shape[2]
        out = torch.empty((n, cin, 2 * ho, 2 * ho), dtype=torch.float32, device=x.device)
    else:
        cin, ho = x.shape[1], x.shape[2] // 2
        out = torch.empty((n, 2 * cin, ho, ho), dtype=torch.float32, device=x.device)
    check(_lib.lib().afan_conv3x3s2_f32(f32(x, "x"), f32(w_packed, "w_packed"), f32(out), n, cin, ho, int(bool(dgrad)), stream()),
          "afan_conv3x3s2_f32")
    return out


def conv3x3s2_wgrad(x: torch.Tensor, dy: torch.Tensor, ws: torch.Tensor, accumulate_into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dW [2C, C, 3, 3] of the stride-2 transition (x [N, C, H, H], dy [N, 2C, H/2, H/2]); deterministic."""
    n, cin, hin, _ = x.shape
    dw = torch.empty((2 * cin, cin, 3, 3), dtype=torch.float32, device=x.device) if accumulate_into is None else accumulate_into
    if dw.numel() != 18 * cin * cin:
        raise AfanError("accumulate_into must hold 2C*C*9 floats")
    check(_lib.lib().afan_conv3x3s2_wgrad_f32(f32(x, "x"), f32(dy, "dy"), f32(dw), ptr(ws), ws.numel() * 4, n, cin, hin // 2,
                                              int(accumulate_into is not None), stream()), "afan_conv3x3s2_wgrad_f32")
    return dw


def conv3x3_wgrad_workspace(c: int, device) -> torch.Tensor:
    return torch.empty(_lib.lib().afan_conv3x3_wgrad_workspace_bytes(int(c)) // 4, dtype=torch.float32, device=device)


def conv3x3_wgrad_umma_supported(n: int, c: int, h: int) -> bool:
    return bool(_lib.lib().afan_conv3x3_wgrad_umma_supported(int(n), int(c), int(h)))


def conv3x3_wgrad(x: torch.Tensor, dy: torch.Tensor, ws: torch.Tensor, accumulate_into: Optional[torch.Tensor] = None,
                  math: str = "fp32") -> torch.Tensor:
    """dW [C, C, 3, 3] of the convolution above (deterministic: per-CTA partials folded in a fixed order).
    accumulate_into: add the result to this tensor (a parameter's .grad inside the gradient arena) instead.
    math: "fp32" (FFMA kernel) or "umma" (tcgen05 3xTF32, (C, H) in {(32, 16), (64, 8)})."""
    n, c, h, _ = x.shape
    dw = torch.empty((c, c, 3, 3), dtype=torch.float32, device=x.device) if accumulate_into is None else accumulate_into
    if dw.numel() != c * c * 9:
        raise AfanError("accumulate_into must hold C*C*9 floats")
    if math == "umma":
        check(_lib.lib().afan_conv3x3_wgrad_umma_f32(f32(x, "x"), f32(dy, "dy"), f32(dw), ptr(ws), ws.numel() * 4, n, c, h,
                                                     int(accumulate_into is not None), stream()), "afan_conv3x3_wgrad_umma_f32")
        return dw
    check(_lib.lib().afan_conv3x3_wgrad_f32(f32(x, "x"), f32(dy, "dy"), f32(dw), ptr(ws), ws.numel() * 4, n, c, h,
                                            int(accumulate_into is not None), stream()), "afan_conv3x3_wgrad_f32")
    return dw
