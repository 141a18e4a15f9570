"""ctypes binding of ``libpetsyn.so`` (the C ABI declared in ``include/petsyn.h``).

This is the *only* place the package touches native code.  There is no CPU implementation and no
alternative backend: if the shared library is missing the import fails loudly, and every kernel entry
point raises when called without a CUDA device.

Error convention (mirrors the reference's Python exceptions): ``PETSYN_EINVAL`` ->
``ValueError``; every other failure -> ``RuntimeError``; the message is ``petsyn_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpetsyn.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C <package>/csrc`). There is no CPU/PyTorch fallback for this package.")

lib = C.CDLL(LIB_PATH)

# ---------------------------------------------------------------------------------------------- constants
OP_CONV, OP_UPCONV, OP_CONVT = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SILU, ACT_TANH, ACT_PRELU = 0, 1, 2, 3, 4, 5
E_INVAL, E_CUDA, E_NOMEM = -1, -2, -3


class ConvDesc(C.Structure):
    """``petsyn_conv_desc`` (include/petsyn.h)."""
    _fields_ = [
        ("op", C.c_int32),
        ("n", C.c_int32), ("d", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("x_cstride", C.c_int32), ("x_coff", C.c_int32),
        ("y_cstride", C.c_int32), ("y_coff", C.c_int32),
        ("dy_cstride", C.c_int32), ("dy_coff", C.c_int32),
        ("dx_cstride", C.c_int32), ("dx_coff", C.c_int32),
        ("epi_act", C.c_int32),
        ("epi_slope", C.c_float),
        ("y_fp32", C.c_int32),
    ]


class NormActDesc(C.Structure):
    """``petsyn_normact_desc`` (include/petsyn.h)."""
    _fields_ = [
        ("z", C.c_void_p), ("rows", C.c_int64), ("c", C.c_int32), ("nsamples", C.c_int32),
        ("per_sample_stats", C.c_int32),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p), ("gamma", C.c_void_p),
        ("t1", C.c_void_p), ("t1_cstride", C.c_int32), ("t1_coff", C.c_int32), ("act1", C.c_int32),
        ("t2", C.c_void_p), ("t2_cstride", C.c_int32), ("t2_coff", C.c_int32), ("act2", C.c_int32),
        ("slope", C.c_float),
        ("res", C.c_void_p), ("res_cstride", C.c_int32), ("res_coff", C.c_int32), ("res_accumulate", C.c_int32),
        ("sums", C.c_void_p), ("dz", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
        ("group_size", C.c_int32), ("dz_accumulate", C.c_int32), ("affine_accumulate", C.c_int32),
        ("slope_dev", C.c_void_p), ("dslope", C.c_void_p), ("dz_colsum", C.c_void_p),
        ("t1_stats", C.c_void_p), ("t1_stats_c", C.c_int32), ("t1_stats_coff", C.c_int32),
        ("t2_stats", C.c_void_p), ("t2_stats_c", C.c_int32), ("t2_stats_coff", C.c_int32),
        ("fin_sums", C.c_void_p), ("fin_gamma", C.c_void_p), ("fin_beta", C.c_void_p),
        ("fin_group_size", C.c_int32), ("fin_eps", C.c_float), ("separate_group_combine", C.c_int32),
        ("sums_prezeroed", C.c_int32),
        ("z_cstride", C.c_int32), ("z_coff", C.c_int32), ("dz_cstride", C.c_int32), ("dz_coff", C.c_int32),
        ("extra", C.c_void_p), ("extra_cstride", C.c_int32), ("extra_coff", C.c_int32),
        ("dz_colsum_coff", C.c_int32), ("dz_colsum_c", C.c_int32),
        ("sums_precomputed", C.c_int32),
    ]


class ConvEpilogue(C.Structure):
    """``petsyn_conv_epilogue`` (include/petsyn.h)."""
    _fields_ = [
        ("side", C.c_void_p), ("side_cstride", C.c_int32), ("side_coff", C.c_int32), ("add_side", C.c_int32),
        ("stats1", C.c_void_p), ("stats1_c", C.c_int32), ("stats1_coff", C.c_int32),
        ("stats2", C.c_void_p), ("stats2_c", C.c_int32), ("stats2_coff", C.c_int32),
        ("norm_scale", C.c_void_p), ("norm_shift", C.c_void_p), ("norm_mean", C.c_void_p), ("norm_rstd", C.c_void_p),
        ("norm_act", C.c_int32), ("norm_slope", C.c_float),
        ("bsums", C.c_void_p),
    ]


class VolumeSrc(C.Structure):
    _fields_ = [("data", C.c_void_p), ("d", C.c_int32), ("h", C.c_int32), ("w", C.c_int32)]


_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol declared in include/petsyn.h
SIGNATURES = {
    "petsyn_version": (_i32, []),
    "petsyn_last_error": (C.c_char_p, []),
    "petsyn_launch_count": (C.c_uint64, []),
    "petsyn_conv_plan_create": (_i32, [C.POINTER(ConvDesc), C.POINTER(_vp)]),
    "petsyn_conv_plan_destroy": (None, [_vp]),
    "petsyn_conv_out_dims": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "petsyn_conv_flops": (_i32, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "petsyn_conv_kernel_path": (_i32, [_vp, _i32]),
    "petsyn_conv_packed_fprop_bytes": (_sz, [_vp]),
    "petsyn_conv_packed_dgrad_bytes": (_sz, [_vp]),
    "petsyn_conv_wgrad_scratch_bytes": (_sz, [_vp]),
    "petsyn_conv_workspace_bytes": (_sz, [_vp]),
    "petsyn_conv_set_workspace": (_i32, [_vp, _vp, _sz]),
    "petsyn_conv_pack_weights": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "petsyn_pack_batch_create": (_i32, [_i32, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                                        C.POINTER(_vp)]),
    "petsyn_pack_batch_run": (_i32, [_vp, _vp]),
    "petsyn_pack_batch_destroy": (None, [_vp]),
    "petsyn_conv_fprop": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "petsyn_conv_dgrad": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "petsyn_conv_dgrad_accumulate": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "petsyn_conv_epilogue_supported": (_i32, [_vp, _i32]),
    "petsyn_conv_fprop_epi": (_i32, [_vp, _vp, _vp, _vp, _vp, C.POINTER(ConvEpilogue), _vp]),
    "petsyn_conv_dgrad_epi": (_i32, [_vp, _vp, _vp, _vp, C.POINTER(ConvEpilogue), _vp]),
    "petsyn_conv_wgrad": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "petsyn_conv_wgrad_bias": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp]),
    "petsyn_stem_col2im_k4s2": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "petsyn_concat_latent": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "petsyn_take_channel0": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "petsyn_put_channel0_grad": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "petsyn_norm_stats": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "petsyn_norm_stats_slice": (_i32, [_vp, _i32, _i32, _vp, _i64, _i32, _i32, _vp]),
    "petsyn_norm_finalize": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _f32, _f32, _i32,
                                    _vp]),
    "petsyn_normact_fwd": (_i32, [C.POINTER(NormActDesc), _vp]),
    "petsyn_normact_bwd": (_i32, [C.POINTER(NormActDesc), _vp]),
    "petsyn_add_slice": (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _i64, _i32, _i32, _vp]),
    "petsyn_colsum": (_i32, [_vp, _i32, _i32, _vp, _i64, _i32, _vp]),
    "petsyn_stem_im2col_k4s2": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "petsyn_head_gather_tanh": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "petsyn_head_scatter_bwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "petsyn_bn_stats": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "petsyn_bn_finalize": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _i32, _vp]),
    "petsyn_norm_act_fwd": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _f32, _i64, _i32, _vp]),
    "petsyn_norm_act_bwd_reduce": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _f32,
                                          _vp, _i64, _i32, _vp]),
    "petsyn_norm_act_bwd_apply": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32,
                                         _f32, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "petsyn_l1_loss_fwd_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _vp]),
    "petsyn_ssim_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "petsyn_ssim_fwd_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _i32, _vp]),
    "petsyn_abs_sq_err": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "petsyn_avgpool2_f32": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "petsyn_mse_const_fwd_bwd": (_i32, [_vp, _f32, _vp, _vp, _i64, _f32, _vp]),
    "petsyn_resample2": (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _vp]),
    "petsyn_layernorm_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    "petsyn_layernorm_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "petsyn_geglu_fwd": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "petsyn_geglu_bwd": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "petsyn_attention_fwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]),
    "petsyn_attention_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]),
    "petsyn_covariate_bias_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _vp]),
    "petsyn_covariate_bias_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _vp]),
    "petsyn_kl_fwd_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp]),
    "petsyn_adam_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _i32, _vp, _vp]),
    "petsyn_sumsq": (_i32, [_vp, _vp, _i64, _vp]),
    "petsyn_volume_prepare": (_i32, [C.POINTER(VolumeSrc), _i32, _vp, _i32, _i32, _i32, _vp, _vp]),
    "petsyn_volume_window_offset": (_i32, [_i32, _i32]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)        # AttributeError here == the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    msg = lib.petsyn_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    """Translate a C return code into the reference's exception convention."""
    if rc == 0:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc == E_INVAL:
        raise ValueError(msg)
    raise RuntimeError(msg)


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda() -> None:
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("petsyn: a CUDA device (B200, sm_100a) is required; there is no CPU fallback")
