"""ctypes binding of libafan_b200.so (the C ABI declared in include/afan_b200.h).

There is NO fallback: if the shared library is missing, or a tensor is not a contiguous CUDA
fp32 tensor, the call raises.  Nothing here imports oracle/.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libafan_b200.so")

_vp, _i64, _f32, _int, _u64, _f64 = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_int,
                                     ctypes.c_uint64, ctypes.c_double)

# name -> (restype, argtypes); mirrors include/afan_b200.h one to one
SIGNATURES = {
    "afan_version": (ctypes.c_char_p, []),
    "afan_strerror": (ctypes.c_char_p, [_int]),
    "afan_device_info": (_int, [ctypes.POINTER(_int)] * 3),
    "afan_ffma_probe": (_int, [_vp, _i64, _i64, ctypes.POINTER(_f64), _vp]),
    "afan_hbm_probe": (_int, [_int, _vp, _vp, _i64, _vp, _vp]),
    "afan_pgd_init_noise_f32": (_int, [_vp, _vp, _vp, _i64, _f32, _vp]),
    "afan_pgd_init_philox_f32": (_int, [_vp, _vp, _i64, _f32, _u64, _u64, _vp, _vp]),
    "afan_pgd_norms_workspace_bytes": (_i64, [_i64]),
    "afan_pgd_linf_step_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _f32, _f32, _int, _vp]),
    "afan_tensor_clamp_f32": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "afan_pgd_linf_step_bf16": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _f32, _f32, _int, _vp]),
    "afan_pgd_init_noise_bf16": (_int, [_vp, _vp, _vp, _i64, _f32, _vp]),
    "afan_pgd_init_philox_bf16": (_int, [_vp, _vp, _i64, _f32, _u64, _u64, _vp, _vp]),
    "afan_sample_l2norm_f32": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "afan_pgd_l2_step_f32": (_int, [_vp, _vp, _vp, _i64, _i64, _f32, _f32, _vp]),
    "afan_l2ball_proj_f32": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _vp]),
    "afan_mix_feature_f32": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "afan_sat_mix_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _i64, _i64, _i64, _vp]),
    "afan_shortcut_a_fwd_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp]),
    "afan_shortcut_a_bwd_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp]),
    "afan_linear_wgrad_f32": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "afan_nms_workspace_bytes": (_i64, [_i64]),
    "afan_nms_f32": (_int, [_vp, _vp, _f32, _vp, _vp, _vp, _i64, _i64, _vp]),
    "afan_nms_batched_workspace_bytes": (_i64, [_i64, _i64]),
    "afan_nms_batched_f32": (_int, [_vp, _f32, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "afan_roi_align_fwd_f32": (_int, [_vp, _vp, _vp] + [_i64] * 7 + [_f32, _int, _vp]),
    "afan_roi_align_bwd_f32": (_int, [_vp, _vp, _vp] + [_i64] * 7 + [_f32, _int, _vp]),
    "afan_bn_workspace_bytes": (_i64, [_i64, _i64]),
    "afan_bn_fwd_f32": (_int, [_vp] * 9 + [_vp, _i64] + [_i64] * 4 + [_f32, _f32, _int, _int, _vp]),
    "afan_bn_bwd_f32": (_int, [_vp] * 10 + [_vp, _i64] + [_i64] * 4 + [_int, _vp]),
    "afan_bn_fwd_stats_f32": (_int, [_vp, _vp, _vp, _i64] + [_i64] * 4 + [_vp]),
    "afan_bn_fwd_finalize_f32": (_int, [_vp, _f64] + [_vp] * 6 + [_vp, _i64, _i64, _i64, _f32, _f32, _int, _vp]),
    "afan_bn_fwd_apply_f32": (_int, [_vp, _vp, _vp, _vp, _i64] + [_i64] * 4 + [_int, _vp]),
    "afan_bn_bwd_reduce_f32": (_int, [_vp] * 8 + [_vp, _i64] + [_i64] * 4 + [_int, _vp]),
    "afan_bn_bwd_finalize_f32": (_int, [_vp, _f64, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "afan_bn_bwd_apply_f32": (_int, [_vp] * 5 + [_vp, _i64] + [_i64] * 4 + [_int, _vp]),
    "afan_bn_mailbox_bytes": (_i64, [_int, _i64]),
    "afan_bn_fwd_p2p_f32": (_int, [_vp] * 9 + [_i64] * 4 + [_f32, _f32, _int, _int, _int, _int, _vp, _i64, _vp, _vp]),
    "afan_bn_bwd_p2p_f32": (_int, [_vp] * 10 + [_i64] * 4 + [_int, _int, _int, _vp, _i64, _vp, _vp]),
    "afan_p2p_alloc": (_int, [ctypes.POINTER(_vp), _i64]),
    "afan_p2p_free": (_int, [_vp]),
    "afan_p2p_get_handle": (_int, [_vp, _vp]),
    "afan_p2p_open_handle": (_int, [_vp, ctypes.POINTER(_vp)]),
    "afan_p2p_close_handle": (_int, [_vp]),
    "afan_bn_affine_f32": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "afan_bn_affine_bwd_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "afan_sgd_momentum_f32": (_int, [_vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _vp]),
    "afan_conv3x3_pack_f32": (_int, [_vp, _i64, _i64, _vp]),
    "afan_conv3x3_f32": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "afan_conv3x3s2_f32": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "afan_conv3x3s2_wgrad_f32": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _vp]),
    "afan_conv3x3_pack_tc_f32": (_int, [_vp, _i64, _i64, _int, _vp]),
    "afan_conv3x3_tc_f32": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _vp]),
    "afan_conv3x3_umma_supported": (_int, [_i64, _i64, _i64]),
    "afan_conv3x3_pack_umma_f32": (_int, [_vp, _i64, _i64, _vp]),
    "afan_conv3x3_umma_f32": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "afan_conv3x3_umma_bn_workspace_bytes": (_i64, [_i64, _i64]),
    "afan_conv3x3_umma_bn_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64,
                                        _f32, _f32, _int, _vp]),
    "afan_bn_bwd_xmask_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "afan_conv3x3_umma_bn_p2p_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _f32,
                                            _int, _int, _int, _vp, _i64, _vp, _vp]),
    "afan_bn_bwd_xmask_p2p_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _vp, _i64,
                                         _vp, _vp]),
    "afan_conv3x3_wgrad_umma_supported": (_int, [_i64, _i64, _i64]),
    "afan_conv3x3_wgrad_umma_f32": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _vp]),
    "afan_conv3x3_wgrad_workspace_bytes": (_i64, [_i64]),
    "afan_conv3x3_wgrad_f32": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _vp]),
}

AFAN_ERR_UNSUPPORTED = -5
_lib = None
# kernels launched per C-ABI call (for bench.py's `gpu_launches` claim); bumped by check() on success
KERNELS_PER_CALL = {"afan_nms_f32": 2, "afan_nms_batched_f32": 2, "afan_conv3x3_wgrad_f32": 2, "afan_conv3x3_wgrad_umma_f32": 2, "afan_conv3x3s2_wgrad_f32": 2}
launch_count = 0


class AfanError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AfanError(f"{LIB_PATH} is missing: the sm_100a extension has not been built "
                            f"(make -C {os.path.join(_HERE, 'csrc')}); there is no CPU/PyTorch fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    global launch_count
    launch_count += KERNELS_PER_CALL.get(what, 1)
    if rc != 0:
        raise AfanError(f"{what or 'afan call'} failed: {lib().afan_strerror(int(rc)).decode()} (code {rc})")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise AfanError("afan_b200 kernels need CUDA tensors; there is no CPU fallback")
    if not t.is_contiguous():
        raise AfanError("afan_b200 kernels need contiguous (NCHW) tensors")
    return t.data_ptr()


def f32(t, name="tensor"):
    if t is not None and t.dtype != torch.float32:
        raise AfanError(f"{name} must be float32, got {t.dtype}")
    return ptr(t)


def dev_ptr(t, dtype, name="tensor"):
    if t is not None and t.dtype != dtype:
        raise AfanError(f"{name} must be {dtype}, got {t.dtype}")
    return ptr(t)


def stream():
    """cudaStream_t of torch's current stream (so launches order with torch ops and get graph-captured)."""
    return torch.cuda.current_stream().cuda_stream


def version() -> str:
    return lib().afan_version().decode()
