"""numpy front-end of the CPU oracle (oracle/afan_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg as the CHECKER.  The product package (cv_a-fan_b200/) never imports it.

Every function takes/returns numpy float32 arrays (C-contiguous) and cites the reference
lines restated by the C routine it wraps.  Parity pin: tests/golden/*.npz (generated from
the live reference by oracle/gen_golden.py), checked in tests/test_oracle_golden.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libafan_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i64 = ctypes.c_int64


def build(force: bool = False) -> str:
    """Compile oracle/afan_oracle.c with gcc (seconds)."""
    src = os.path.join(_HERE, "afan_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    if a is None:
        return ctypes.cast(None, _f32p)
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_f32p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def pgd_init_noise(x, u, eps):
    """Classification/attack_algo.py:42-44 (random start with the given torch.rand draw)."""
    x, u = _f32(x), _f32(u)
    out = np.empty_like(x)
    lib().orc_pgd_init_noise_f32(_p(x), _p(u), _p(out), _i64(x.size), ctypes.c_float(eps))
    return out


def pgd_linf_step(grad, x_clean, x_adv, gamma, eps, clip):
    """Classification/attack_algo.py:53-56 (+ :9-19,35-36).  Returns the new x_adv."""
    grad, x_adv = _f32(grad), _f32(x_adv).copy()
    x_clean = _f32(x_clean) if x_clean is not None else None
    assert not clip or x_clean is not None
    lib().orc_pgd_linf_step_f32(_p(grad), _p(x_clean), _p(x_adv), _i64(x_adv.size),
                                ctypes.c_float(gamma), ctypes.c_float(eps), int(bool(clip)))
    return x_adv


def _u16(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint16
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16))


def pgd_linf_step_bf16(grad_u16, x_clean_u16, x_adv_u16, gamma, eps, clip, want_delta=False):
    """bf16-storage twin of pgd_linf_step on uint16 bit patterns (no reference for this dtype: unpinned).
    Returns (x_adv_new_u16, delta_u16 | None)."""
    g, gp = _u16(grad_u16)
    xa, xap = _u16(np.array(x_adv_u16, copy=True))
    xcp = ctypes.cast(None, ctypes.POINTER(ctypes.c_uint16))
    if x_clean_u16 is not None:
        xc, xcp = _u16(x_clean_u16)
    d, dp = (None, ctypes.cast(None, ctypes.POINTER(ctypes.c_uint16)))
    if want_delta:
        d, dp = _u16(np.empty_like(xa))
    lib().orc_pgd_linf_step_bf16(gp, xcp, xap, dp, _i64(xa.size), ctypes.c_float(gamma), ctypes.c_float(eps), int(bool(clip)))
    return xa, d


def pgd_init_noise_bf16(x_u16, u, eps):
    x, xp = _u16(x_u16)
    u = _f32(u)
    out, outp = _u16(np.empty_like(x))
    lib().orc_pgd_init_noise_bf16(xp, _p(u), outp, _i64(x.size), ctypes.c_float(eps))
    return out


def delta_norms(x_adv, x_clean):
    """Classification/main_perturb.py:188-192 -> (delta, l2[N], linf[N])."""
    x_adv, x_clean = _f32(x_adv), _f32(x_clean)
    n = x_adv.shape[0]
    per = x_adv.size // max(n, 1)
    delta = np.empty_like(x_adv)
    l2 = np.empty(n, np.float32)
    linf = np.empty(n, np.float32)
    lib().orc_delta_norms_f32(_p(x_adv), _p(x_clean), _p(delta), _p(l2), _p(linf), _i64(n), _i64(per))
    return delta, l2, linf


def l2ball_proj(center, radius, t):
    """Classification/attack_algo.py:21-33.  Returns the projected copy of t."""
    center, t = _f32(center), _f32(t).copy()
    n = t.shape[0]
    lib().orc_l2ball_proj_f32(_p(center), ctypes.c_float(radius), _p(t), _i64(n), _i64(t.size // max(n, 1)))
    return t


def pgd_l2_step(grad, x_clean, x_adv, gamma, eps, clip, tiny=1e-12):
    """L2-normalised step (a5b; defined by this build, see afan_oracle.c)."""
    grad, x_adv = _f32(grad), _f32(x_adv).copy()
    x_clean = _f32(x_clean) if x_clean is not None else None
    n = x_adv.shape[0]
    lib().orc_pgd_l2_step_f32(_p(grad), _p(x_clean), _p(x_adv), _i64(n), _i64(x_adv.size // max(n, 1)),
                              ctypes.c_float(gamma), ctypes.c_float(eps), int(bool(clip)), ctypes.c_float(tiny))
    return x_adv


def mix_feature(clean, adv):
    """Segmentation/attack_algo.py:121-130 == Detection/attack_algo.py:254-265 (NCHW)."""
    clean, adv = _f32(clean), _f32(adv)
    n, c = clean.shape[0], clean.shape[1]
    hw = clean.size // max(n * c, 1)
    out = np.empty_like(clean)
    lib().orc_mix_feature_f32(_p(clean), _p(adv), _p(out), _i64(n), _i64(c), _i64(hw))
    return out


def lerp(x, y, w):
    """torch.lerp as used by get_sample_points, Segmentation/attack_algo.py:108-118."""
    x, y = _f32(x), _f32(y)
    out = np.empty_like(x)
    lib().orc_lerp_f32(_p(x), _p(y), ctypes.c_float(w), _p(out), _i64(x.size))
    return out


def get_sample_points(x, y, number):
    """Segmentation/attack_algo.py:108-118: [x, lerp(x,y,i/(number-1)) ..., y]."""
    percent = 1.0 / (number - 1)
    pts = [x] + [lerp(x, y, i * percent) for i in range(1, number - 1)] + [y]
    return pts


def bn_fwd(x, weight, bias, running_mean, running_var, groups=1, residual=None, relu=False,
           eps=1e-5, momentum=0.1, replay=1):
    """Train-mode BatchNorm2d over `groups` statistic groups of a [G*N,C,H,W] batch
    (resnet_s.py:54,56,89 applied to the adv and clean batches, main_perturb.py:195-196).
    Returns (y, save_mean[G,C], save_invstd[G,C]); running_* are updated IN PLACE."""
    x = _f32(x)
    gn, c = x.shape[0], x.shape[1]
    hw = x.size // (gn * c)
    n = gn // groups
    y = np.empty_like(x)
    sm = np.empty((groups, c), np.float32)
    si = np.empty((groups, c), np.float32)
    for a in (running_mean, running_var):
        assert a is None or (a.dtype == np.float32 and a.flags["C_CONTIGUOUS"])
    lib().orc_bn_fwd_f32(_p(x), _p(_f32(residual) if residual is not None else None),
                         _p(_f32(weight) if weight is not None else None),
                         _p(_f32(bias) if bias is not None else None),
                         _p(running_mean), _p(running_var), _p(y), _p(sm), _p(si),
                         _i64(groups), _i64(n), _i64(c), _i64(hw), ctypes.c_float(eps),
                         ctypes.c_float(momentum), int(bool(relu)), int(replay))
    return y, sm, si


def bn_bwd(dy, x, y, weight, save_mean, save_invstd, groups=1, relu=False, residual=False):
    """Backward of bn_fwd -> (dx, dresidual|None, dweight[C], dbias[C])."""
    dy, x = _f32(dy), _f32(x)
    gn, c = x.shape[0], x.shape[1]
    hw = x.size // (gn * c)
    n = gn // groups
    dx = np.empty_like(x)
    dres = np.empty_like(x) if residual else None
    dw = np.empty(c, np.float32)
    db = np.empty(c, np.float32)
    lib().orc_bn_bwd_f32(_p(dy), _p(x), _p(_f32(y) if y is not None else None),
                         _p(_f32(weight) if weight is not None else None),
                         _p(_f32(save_mean)), _p(_f32(save_invstd)), _p(dx), _p(dres), _p(dw), _p(db),
                         _i64(groups), _i64(n), _i64(c), _i64(hw), int(bool(relu)))
    return dx, dres, dw, db


def sgd_momentum(p, g, buf, lr, momentum, weight_decay):
    """torch.optim.SGD step (main_perturb.py:72-74,201) -> (p', buf')."""
    p, g, buf = _f32(p).copy(), _f32(g), _f32(buf).copy()
    lib().orc_sgd_momentum_f32(_p(p), _p(g), _p(buf), _i64(p.size), ctypes.c_float(lr),
                               ctypes.c_float(momentum), ctypes.c_float(weight_decay))
    return p, buf


def philox_uniform(n, seed, offset=0):
    """Philox4x32-10 uniform stream of the on-device random start (fast path of a2)."""
    u = np.empty(n, np.float32)
    lib().orc_philox_uniform_f32(_p(u), _i64(n), ctypes.c_uint64(seed), ctypes.c_uint64(offset))
    return u


def nms(boxes, scores, thr, strict_gt=True):
    """Detection/support/src/cuda/nms.cu (IoU > thr) / cpu/nms_cpu.cpp (>=): kept ORIGINAL indices, ascending."""
    boxes = _f32(boxes)
    n = boxes.shape[0]
    order = np.ascontiguousarray(np.argsort(-np.asarray(scores, np.float32), kind="stable"), dtype=np.int64)
    keep = np.zeros(max(n, 1), np.uint8)
    lib().orc_nms_f32.restype = ctypes.c_int64
    lib().orc_nms_f32(_p(boxes), order.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _i64(n), ctypes.c_float(thr),
                      int(bool(strict_gt)), keep.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return np.nonzero(keep[:n])[0].astype(np.int64)


def roi_align(feat, rois, output_size, spatial_scale, sampling_ratio=0, dout=None):
    """Detection/support/src/cuda/ROIAlign_cuda.cu fwd (:64-122) and, when dout is given, bwd (:177-254).
    Returns out [R,C,PH,PW] (and dfeat [N,C,H,W] when dout is given)."""
    feat, rois = _f32(feat), _f32(rois)
    n, c, h, w = feat.shape
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    r = rois.shape[0]
    out = np.zeros((r, c, ph, pw), np.float32)
    args = (_i64(c), _i64(h), _i64(w), _i64(r), _i64(ph), _i64(pw), ctypes.c_float(spatial_scale), int(sampling_ratio))
    lib().orc_roi_align_f32(_p(feat), _p(rois), _p(out), _p(None), _p(None), *args)
    if dout is None:
        return out
    dout = _f32(dout)
    dfeat = np.zeros_like(feat)
    lib().orc_roi_align_f32(_p(None), _p(rois), _p(None), _p(dout), _p(dfeat), *args)
    return out, dfeat
