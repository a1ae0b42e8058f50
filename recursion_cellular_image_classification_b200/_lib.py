"""ctypes binding of librxb.so (include/rxb.h).

There is deliberately no fallback: if the library is missing, was not built for sm_100a, or a call
fails, an exception is raised.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# RXB_LIB (development): an alternative build of the same ABI, for same-box A/B timing
LIB_PATH = os.environ.get("RXB_LIB") or os.path.join(_HERE, "librxb.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_int64 = ctypes.c_int64
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t


class RxbError(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    """struct rxb_conv_desc (include/rxb.h)."""
    _fields_ = [(n, c_int) for n in (
        "B", "H", "W", "Cin", "ldA", "Cout", "ldC", "c_off", "taps_y", "taps_x", "pad_y", "pad_x",
        "prologue", "stats")]


class Dn121Config(ctypes.Structure):
    """struct rxb_dn121_config (include/rxb.h)."""
    _fields_ = [("B", c_int), ("H", c_int), ("W", c_int), ("num_classes", c_int),
                ("bn_eps", c_float), ("bn_momentum", c_float)]


class Rn50Config(ctypes.Structure):
    """struct rxb_rn50_config (include/rxb.h)."""
    _fields_ = [("B", c_int), ("G", c_int), ("H", c_int), ("W", c_int), ("num_classes", c_int),
                ("size_features", c_int), ("bn_eps", c_float), ("bn_momentum", c_float)]


# name -> (restype, argtypes); every symbol include/rxb.h declares
SIGNATURES = {
    "rxb_version": (c_int, []),
    "rxb_last_error": (ctypes.c_char_p, []),
    "rxb_check_device": (c_int, []),
    "rxb_stats_accumulate": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "rxb_stats_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "rxb_load_norm_aug": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rxb_jpeg_decode_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "rxb_jpeg_decode_gray": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "rxb_load_norm_affine": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rxb_tta_softmax_avg_mask": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rxb_mask_rescale": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rxb_greedy_assign_workspace_bytes": (c_size_t, [c_int, c_int]),
    "rxb_greedy_assign": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rxb_softmax_ce": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p]),
    "rxb_sgd_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_int, c_float,
                             c_void_p]),
    "rxb_conv_fwd": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p]),
    "rxb_conv_dgrad_bn": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                  c_int, c_void_p, c_void_p, c_void_p]),
    "rxb_conv_dgrad_bn_ex": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rxb_conv_dgrad_bn_wgrad": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p]),
    "rxb_conv_dgrad3x3_bn_wgrad_fixup": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_int, c_int, c_void_p,
                                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                                 c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rxb_bn_sum_dyx_from_wdw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "rxb_conv_wgrad": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                               c_void_p, c_void_p]),
    "rxb_dn121_param_count": (c_int64, [ctypes.POINTER(Dn121Config)]),
    "rxb_dn121_buffer_count": (c_int64, [ctypes.POINTER(Dn121Config)]),
    "rxb_dn121_workspace_bytes": (c_size_t, [ctypes.POINTER(Dn121Config), c_int]),
    "rxb_dn121_create": (c_int, [ctypes.POINTER(Dn121Config), c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_size_t, c_int, ctypes.POINTER(c_void_p)]),
    "rxb_dn121_destroy": (None, [c_void_p]),
    "rxb_dn121_sync_weights": (c_int, [c_void_p, c_void_p]),
    "rxb_dn121_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "rxb_dn121_num_phases": (c_int, []),
    "rxb_dn121_phase_grad_range": (c_int, [c_void_p, c_int, ctypes.POINTER(c_int64), ctypes.POINTER(c_int64)]),
    "rxb_dn121_train_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "rxb_dn121_sgd": (c_int, [c_void_p, c_float, c_float, c_float, c_int, c_float, c_void_p]),
    "rxb_rn50_param_count": (c_int64, [ctypes.POINTER(Rn50Config)]),
    "rxb_rn50_buffer_count": (c_int64, [ctypes.POINTER(Rn50Config)]),
    "rxb_rn50_head_offset": (c_int64, [ctypes.POINTER(Rn50Config)]),
    "rxb_rn50_workspace_bytes": (c_size_t, [ctypes.POINTER(Rn50Config), c_int]),
    "rxb_rn50_create": (c_int, [ctypes.POINTER(Rn50Config), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                c_int, ctypes.POINTER(c_void_p)]),
    "rxb_rn50_train_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "rxb_rn50_sgd": (c_int, [c_void_p, c_float, c_float, c_float, c_int, c_float, c_int, c_void_p]),
    "rxb_rn50_destroy": (None, [c_void_p]),
    "rxb_rn50_sync_weights": (c_int, [c_void_p, c_void_p]),
    "rxb_rn50_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "rxb_launch_count": (c_int64, []),
    "rxb_launch_count_reset": (None, []),
    "rxb_profile_enable": (None, [c_int]),
    "rxb_profile_collect": (c_int, [ctypes.POINTER(c_float), ctypes.POINTER(c_int64), c_int]),
}

_lib = None


def load():
    """Load librxb.so and bind every declared symbol.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RxbError(
            "%s not found: build it with `python -m recursion_cellular_image_classification_b200.build` "
            "(needs nvcc). There is no CPU/PyTorch fallback for the hot path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().rxb_last_error()
        raise RxbError("librxb error %d: %s" % (rc, (msg or b"").decode("utf-8", "replace")))


def require_gpu():
    """Fail loudly when no B200 is visible — the product path never runs on the CPU."""
    if not torch.cuda.is_available():
        raise RxbError("no CUDA device visible: librxb kernels are sm_100a-only and have no CPU fallback")
    check(load().rxb_check_device())


def ptr(t):
    """Device (or host) address of a contiguous tensor, or None."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise RxbError("tensor passed to librxb must be contiguous")
    return c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)
