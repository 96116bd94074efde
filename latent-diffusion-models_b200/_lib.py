"""ctypes binding of libldm_b200.so (the C ABI declared in include/ldm_b200.h).

There is no CPU fallback: every entry point raises if the CUDA library is missing or a call fails.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libldm_b200.so")
_lib: Optional[C.CDLL] = None

F32, BF16 = 0, 1
DTYPES = {"fp32": F32, "f32": F32, "float32": F32, "bf16": BF16, "bfloat16": BF16}

c_i64p = C.POINTER(C.c_int64)
vp = C.c_void_p


class UNetDesc(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("out_channels", C.c_int32), ("channels", C.c_int32),
                ("n_levels", C.c_int32), ("channel_multipliers", C.c_int32 * 8),
                ("with_time_emb", C.c_int32), ("num_classes", C.c_int32), ("image_size", C.c_int32),
                ("dtype", C.c_int32), ("conv_impl", C.c_int32)]


class SamplerDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("n_steps", C.c_int32), ("cfg_scale", C.c_float),
                ("y_len", C.c_int32), ("use_graph", C.c_int32)]


class ProfileFamily(C.Structure):
    _fields_ = [("ms", C.c_double), ("flops", C.c_double), ("bytes", C.c_double), ("launches", C.c_int64)]


class Profile(C.Structure):
    _fields_ = [("family", ProfileFamily * 8)]


FAMILIES = ["conv_tc", "conv_ffma", "group_norm", "linear_attention", "attention", "other", "conv_tc_1x1", "conv_halo"]

# name -> (restype, argtypes); mirrors include/ldm_b200.h one to one
SIGNATURES = {
    "ldm_abi_version": (C.c_int, []),
    "ldm_last_error": (C.c_char_p, []),
    "ldm_launch_count": (C.c_int64, []),
    "ldm_reset_launch_count": (None, []),
    "ldm_unet_create": (C.c_int, [C.POINTER(UNetDesc), C.POINTER(vp)]),
    "ldm_unet_destroy": (None, [vp]),
    "ldm_unet_num_params": (C.c_int, [vp]),
    "ldm_unet_param_name": (C.c_char_p, [vp, C.c_int]),
    "ldm_unet_param_numel": (C.c_int64, [vp, C.c_int]),
    "ldm_unet_load_params": (C.c_int, [vp, C.POINTER(vp), C.c_int, vp]),
    "ldm_unet_workspace_bytes": (C.c_int64, [vp, C.c_int]),
    "ldm_unet_forward": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int64, vp]),
    "ldm_unet_profile": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int64, vp, C.POINTER(Profile)]),
    "ldm_unet_set_tap": (C.c_int, [vp, C.c_char_p, vp, C.c_int64]),
    "ldm_q_sample": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int64, C.c_uint64, C.c_uint64, vp]),
    "ldm_p_sample": (C.c_int, [vp, vp, vp, C.c_float, vp, C.c_int, vp, C.c_int, vp, C.c_uint64, C.c_uint64, vp,
                               C.c_int, C.c_int64, vp]),
    "ldm_build_coef_table": (C.c_int, [vp, vp, vp, C.c_int, vp, vp]),
    "ldm_randn": (C.c_int, [vp, C.c_int, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint64, vp]),
    "ldm_sampler_create": (C.c_int, [vp, C.POINTER(SamplerDesc), C.POINTER(vp)]),
    "ldm_sampler_destroy": (None, [vp]),
    "ldm_sampler_workspace_bytes": (C.c_int64, [vp]),
    "ldm_sampler_run": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int, vp,
                                  C.c_int64, vp]),
    "ldm_group_norm": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_float, C.c_int, C.c_int, vp, vp]),
    "ldm_group_norm_workspace_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "ldm_conv2d": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, C.c_int, vp,
                             C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_conv2d_gn_scratch_bytes": (C.c_int64, [C.c_int]),
    "ldm_conv2d_gn": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, vp, vp, vp, C.c_int, vp, C.c_int,
                                C.c_int, C.c_int, vp, C.c_int64, C.c_uint, C.POINTER(C.c_int), vp]),
    "ldm_pack_conv_weight": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp]),
    "ldm_conv_transpose2x2": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, vp]),
    "ldm_pack_conv_transpose_weight": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, vp]),
    "ldm_max_pool2x2": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_linear_attention": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_attention": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    # ---- training step
    "ldm_conv2d_wgrad": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, vp]),
    "ldm_conv2d_wgrad_scratch_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ldm_conv2d_wgrad_tc": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                      vp, vp]),
    "ldm_pack_conv_weight_pair": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp]),
    "ldm_pack_conv_weight_dgrad": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]),
    "ldm_column_sum": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_group_norm_rowvec": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, vp, vp]),
    "ldm_group_norm_backward_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ldm_group_norm_backward": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, vp, vp, vp]),
    "ldm_max_pool2x2_backward": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, vp]),
    "ldm_pixel_unshuffle2x2": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_linear_attention_backward_workspace_bytes": (C.c_int64, [C.c_int]),
    "ldm_linear_attention_backward": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
    "ldm_attention_backward": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_initial_conv": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "ldm_initial_conv_wgrad_scratch_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ldm_initial_conv_wgrad": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "ldm_final_conv": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_final_conv_backward": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_time_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "ldm_time_embed": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]),
    "ldm_time_embed_backward": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]),
    "ldm_time_proj": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
    "ldm_time_proj_backward": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
    "ldm_add": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int, vp]),
    "ldm_copy_channels": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int64, C.c_int, vp]),
    "ldm_linear_attention_qkv": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_linear_attention_prenorm_scratch_bytes": (C.c_int64, [C.c_int]),
    "ldm_linear_attention_prenorm": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, C.c_float, vp, C.c_int, C.c_int, C.c_int, vp,
                                               C.c_int64, vp]),
    "ldm_linear_attention_prenorm_to_out": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, C.c_float, vp, vp, vp, C.c_int, vp, C.c_int,
                                                      C.c_int, vp, C.c_int64, vp]),
    "ldm_upsample_nearest2x": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_downsample_pick": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_attention_single_head": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_gaussian_distribution": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_adam_step": (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, vp]),
    "ldm_images_to_uint8": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_mse": (C.c_int, [vp, vp, vp, C.c_int64, vp]),
    "ldm_mse_backward": (C.c_int, [vp, vp, vp, vp, C.c_int64, vp]),
    "ldm_nchw_to_nhwc": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_nhwc_to_nchw": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
}


class LdmError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library; raises (no fallback) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LdmError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` to build the sm_100a kernels; "
                       "there is no CPU or PyTorch fallback for this path")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.ldm_abi_version() != 1:
        raise LdmError("libldm_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise LdmError(load().ldm_last_error().decode(errors="replace"))


def launch_count() -> int:
    return int(load().ldm_launch_count())


def reset_launch_count() -> None:
    load().ldm_reset_launch_count()


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


# ---- native handle destruction that is safe next to CUDA-graph capture
# ldm_unet_destroy frees the packed-weight arena with cudaFree and ldm_sampler_destroy destroys graph executables: both are
# "unsafe" calls while ANY stream of the process is capturing (global capture mode) -- they invalidate the capture.  Objects
# die whenever Python's cyclic collector runs, e.g. in the middle of torch.cuda.make_graphed_callables of a later model, so a
# destroy request that arrives during a capture is parked and carried out by the next request that arrives outside one.
_graveyard: list = []


def _capturing() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available() and torch.cuda.is_current_stream_capturing())
    except Exception:
        return False


def destroy_native(fn_name: str, handle) -> None:
    """Call ``lib.<fn_name>(handle)`` now, or after the stream capture that is in progress on this thread."""
    if handle is None:
        return
    _graveyard.append((fn_name, handle))
    flush_graveyard()


def flush_graveyard() -> None:
    if not _graveyard or _capturing():
        return
    try:
        lib = load()
    except Exception:
        return
    while _graveyard:
        name, h = _graveyard.pop()
        try:
            getattr(lib, name)(h)
        except Exception:
            pass
