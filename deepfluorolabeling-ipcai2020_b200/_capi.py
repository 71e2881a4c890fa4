"""ctypes binding of include/fluoro_unet.h (the C ABI of libfluorounet.so)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLUORO_UNET_LIB") or os.path.join(HERE, "libfluorounet.so")   # (override: A/B builds of the same sources)

FU_OK = 0
FU_ERR_INVALID_CONFIG = -1
FU_ERR_UNSUPPORTED_SHAPE = -2
# include/fluoro_unet.h FU_PRECISION_*: "parity_tc" = fp32 storage with split-bf16 x3 tensor-core contractions
PRECISION = {"fp32": 0, "bf16": 1, "parity_tc": 2}


class FuConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "in_channels", "n_classes", "depth", "wf", "padding", "pad_mode_zeros", "batch_norm",
        "up_mode_upconv", "max_pool", "num_lands", "do_res", "block_depth", "lands_block_depth",
        "lands_num_1x1", "do_soft_max", "precision")]


class FuTensorInfo(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("ndim", C.c_int32), ("shape", C.c_int64 * 4),
                ("kind", C.c_int32), ("dtype", C.c_int32), ("grad_offset", C.c_int64)]


class FuCounters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "kernel_launches", "tc_kernel_launches", "forward_calls", "backward_calls", "arena_bytes",
        "last_fwd_launches", "last_bwd_launches")]


class FuLossDesc(C.Structure):
    """include/fluoro_unet.h: fu_loss_desc"""
    _fields_ = [("seg", C.c_void_p), ("seg_stride", C.c_int64 * 3),
                ("mask", C.c_void_p), ("mask_stride", C.c_int64 * 3),
                ("heat", C.c_void_p), ("heat_stride", C.c_int64 * 3),
                ("heat_t", C.c_void_p), ("heat_t_stride", C.c_int64 * 3),
                ("B", C.c_int32), ("n_classes", C.c_int32), ("num_lands", C.c_int32),
                ("Ht", C.c_int32), ("Wt", C.c_int32), ("skip_bg", C.c_int32),
                ("dice_wgt", C.c_float), ("heat_wgt", C.c_float)]


# include/fluoro_unet.h: fu_bucket_callback(user, bucket, offset, numel)
BUCKET_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int64, C.c_int64)

EXPORTS = ["fu_engine_create", "fu_engine_destroy", "fu_last_error", "fu_num_tensors",
           "fu_tensor_get_info", "fu_grad_numel", "fu_bind_tensors", "fu_forward", "fu_backward",
           "fu_set_bucket_callback", "fu_early_grad_numel",
           "fu_get_counters", "fu_build_info", "fu_test_conv", "fu_profile_enable", "fu_profile_report", "fu_debug_copy",
           "fu_loss_workspace_doubles", "fu_loss_forward", "fu_loss_backward", "fu_forward_loss", "fu_backward_loss",
           "fu_prep_tiles", "fu_heatmap_targets", "fu_ensemble_workspace_words", "fu_ensemble_combine",
           "fu_extract_landmarks"]

_lib = None


def lib():
    """Load the library; there is deliberately no fallback when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.fu_engine_create.argtypes = [C.POINTER(FuConfig), i32, C.POINTER(vp)]
    L.fu_engine_create.restype = i32
    L.fu_engine_destroy.argtypes = [vp]
    L.fu_engine_destroy.restype = None
    L.fu_last_error.argtypes = [vp]
    L.fu_last_error.restype = C.c_char_p
    L.fu_num_tensors.argtypes = [vp]
    L.fu_num_tensors.restype = i32
    L.fu_tensor_get_info.argtypes = [vp, i32, C.POINTER(FuTensorInfo)]
    L.fu_tensor_get_info.restype = i32
    L.fu_grad_numel.argtypes = [vp]
    L.fu_grad_numel.restype = i64
    L.fu_bind_tensors.argtypes = [vp, C.POINTER(vp), i32]
    L.fu_bind_tensors.restype = i32
    L.fu_forward.argtypes = [vp, vp, i32, i32, i32, i32, i32, i64, vp, vp, vp, vp]
    L.fu_forward.restype = i32
    L.fu_backward.argtypes = [vp, vp, vp, vp, vp]
    L.fu_backward.restype = i32
    L.fu_set_bucket_callback.argtypes = [vp, BUCKET_CALLBACK, vp, vp]
    L.fu_set_bucket_callback.restype = i32
    L.fu_early_grad_numel.argtypes = [vp]
    L.fu_early_grad_numel.restype = i64
    L.fu_get_counters.argtypes = [vp, C.POINTER(FuCounters)]
    L.fu_get_counters.restype = i32
    L.fu_profile_enable.argtypes = [vp, i32]
    L.fu_profile_enable.restype = i32
    L.fu_profile_report.argtypes = [vp, C.c_char_p, i64]
    L.fu_profile_report.restype = i64
    L.fu_debug_copy.argtypes = [vp, C.c_char_p, vp, i64, C.POINTER(C.c_int32)]
    L.fu_debug_copy.restype = i32
    L.fu_build_info.argtypes = []
    L.fu_build_info.restype = C.c_char_p
    L.fu_test_conv.argtypes = [i32] * 12 + [vp] * 8
    L.fu_test_conv.restype = i32
    L.fu_loss_workspace_doubles.argtypes = [i32, i32, i32]
    L.fu_loss_workspace_doubles.restype = i64
    L.fu_loss_forward.argtypes = [C.POINTER(FuLossDesc), vp, vp, vp]
    L.fu_loss_forward.restype = i32
    L.fu_loss_backward.argtypes = [C.POINTER(FuLossDesc), vp, vp, i32, i32, i32, i32, vp, vp, vp]
    L.fu_loss_backward.restype = i32
    L.fu_forward_loss.argtypes = [vp, vp, i32, i32, i32, i64, C.POINTER(FuLossDesc), i32, i32, vp, vp, vp, vp, vp]
    L.fu_forward_loss.restype = i32
    L.fu_backward_loss.argtypes = [vp, C.POINTER(FuLossDesc), i32, i32, vp, vp, vp, vp, vp]
    L.fu_backward_loss.restype = i32
    f32 = C.c_float
    L.fu_prep_tiles.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp]
    L.fu_prep_tiles.restype = i32
    L.fu_heatmap_targets.argtypes = [vp, i32, i32, i32, i32, f32, vp, vp]
    L.fu_heatmap_targets.restype = i32
    L.fu_ensemble_workspace_words.argtypes = [i32, i32]
    L.fu_ensemble_workspace_words.restype = i64
    L.fu_ensemble_combine.argtypes = [C.POINTER(vp), C.POINTER(vp)] + [i32] * 10 + [vp, vp, vp, vp]
    L.fu_ensemble_combine.restype = i32
    L.fu_extract_landmarks.argtypes = [vp, vp, C.POINTER(C.c_int32), i32, i32, i32, i32, i32, f32, f32, vp, vp, vp]
    L.fu_extract_landmarks.restype = i32
    _lib = L
    return L


def last_error(handle=None):
    msg = lib().fu_last_error(handle)
    return msg.decode() if msg else ""
