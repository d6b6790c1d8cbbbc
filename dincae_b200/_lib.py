"""ctypes binding of the C ABI declared in include/dincae_b200.h.

The shared library is built in-tree by ``python -m dincae_b200.build`` (nvcc, sm_100a).
There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_uint16, c_uint64, c_uint8, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DINCAE_LIB") or os.path.join(_HERE, "libdincae_b200.so")   # DINCAE_LIB: A/B builds

_lib = None

vp = c_void_p

# name -> (restype, argtypes); mirrors include/dincae_b200.h one to one
SIGNATURES = {
    "dincae_version": (c_char_p, []),
    "dincae_error_string": (c_char_p, [c_int]),
    "dincae_last_cuda_error": (c_int, []),
    "dincae_launch_count": (ctypes.c_longlong, [c_int]),
    "dincae_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "dincae_prepare_cube": (c_int, [vp, vp, c_int, c_int, c_int, c_double, c_int, vp, vp, vp, vp]),
    "dincae_build_batch": (c_int, [vp, vp, vp, vp, vp, vp, vp, POINTER(c_double), POINTER(c_double),
                                   c_int, c_int, c_int, c_int, c_int, vp, vp, c_int, vp, c_uint64,
                                   vp, vp, c_int, vp, vp, vp]),
    "dincae_build_batch_p8": (c_int, [vp, vp, vp, vp, vp, vp, vp, POINTER(c_double), POINTER(c_double),
                                      c_int, c_int, c_int, c_int, c_int, vp, vp, c_int, vp, c_uint64,
                                      vp, vp, c_int, vp, c_int, vp, vp, vp]),
    "dincae_set_conv_backend": (c_int, [c_int]),
    "dincae_get_conv_backend": (c_int, []),
    "dincae_conv3x3_fwd": (c_int, [vp, c_int, c_int, vp, c_int, vp, vp, c_int, c_int, c_int, c_int,
                                   c_int, c_float, vp, vp, vp, vp]),
    "dincae_conv3x3_dgrad": (c_int, [vp, vp, vp, c_float, c_int, vp, c_int, c_int, c_int, c_int,
                                     c_int, c_int, vp, c_float, vp, c_int, vp, vp, vp]),
    "dincae_conv3x3_wgrad_workspace": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "dincae_conv3x3_wgrad_on_tensor_cores": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "dincae_conv3x3_wgrad": (c_int, [vp, c_int, c_int, vp, c_int, vp, vp, vp, c_float, c_int,
                                     c_int, c_int, c_int, c_int, vp, vp, vp, c_size_t, vp]),
    "dincae_conv3x3_wgrad_partial": (c_int, [vp, c_int, c_int, vp, c_int, vp, vp, vp, c_float, c_int,
                                             c_int, c_int, c_int, c_int, vp, c_size_t, POINTER(c_int), vp]),
    "dincae_conv3x3_wgrad_reduce": (c_int, [vp, c_int, c_int, c_int, c_int, vp, vp, vp]),
    "dincae_conv_weights_bf16_elems": (c_size_t, [c_int, c_int, c_int, POINTER(c_size_t)]),
    "dincae_conv_weights_bf16": (c_int, [vp, c_int, c_int, c_int, c_int, vp, vp, vp]),
    "dincae_p4_to_p8": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, vp, vp]),
    "dincae_p8_to_p4": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, vp, vp]),
    "dincae_conv3x3_fwd_bf16": (c_int, [vp, c_int, c_int, vp, c_int, vp, vp, c_int, c_int, c_int, c_int,
                                        c_int, c_float, vp, vp, vp, vp, c_int, vp]),
    "dincae_conv3x3_dgrad_bf16": (c_int, [vp, vp, vp, c_float, c_int, vp, c_int, c_int, c_int, c_int,
                                          c_int, vp, c_float, vp, c_int, vp, vp, vp]),
    "dincae_conv3x3_wgrad_bf16_workspace": (c_size_t, [c_int, c_int]),
    "dincae_conv3x3_wgrad_bf16": (c_int, [vp, c_int, c_int, vp, c_int, vp, vp, vp, c_float, c_int,
                                          c_int, c_int, c_int, c_int, c_int, vp, vp, vp, c_size_t, vp]),
    "dincae_head_loss_bf16": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_float, vp, vp, vp, vp, vp,
                                      c_size_t, vp]),
    "dincae_dense_workspace": (c_size_t, [c_int, c_int, c_int]),
    "dincae_dense_fwd": (c_int, [vp, vp, vp, c_int, c_int, c_int, c_int, vp, vp, c_size_t, vp]),
    "dincae_dense_bwd_data": (c_int, [vp, vp, vp, c_int, c_int, c_int, vp, vp, c_size_t, vp]),
    "dincae_dense_bwd_weight": (c_int, [vp, vp, c_int, c_int, c_int, vp, vp, vp]),
    "dincae_dense_tc_weights_elems": (c_size_t, [c_int, c_int]),
    "dincae_dense_tc_tile_weights": (c_int, [vp, c_int, c_int, c_int, vp, vp]),
    "dincae_dense_tc_workspace": (c_size_t, [c_int, c_int, c_int]),
    "dincae_dense_tc_fwd": (c_int, [vp, c_int, vp, vp, c_int, c_int, c_int, c_int, vp, c_int, vp, c_size_t, vp]),
    "dincae_dense_tc_bwd_data": (c_int, [vp, c_int, vp, vp, c_int, c_int, c_int, vp, c_int, vp, c_size_t, vp]),
    "dincae_dropout_fwd": (c_int, [vp, c_int, c_int, c_int, c_float, c_uint64, c_uint64, c_int, vp, vp, vp]),
    "dincae_scale_rows": (c_int, [vp, c_int, c_int, c_int, c_float, vp]),
    "dincae_dense_factored_workspace": (c_size_t, [c_int, c_int, c_int]),
    "dincae_dense_factored_sumsq": (c_int, [vp, c_size_t, c_int, vp, c_size_t, c_int, c_int, c_int, c_int, c_int,
                                            vp, vp, c_size_t, vp]),
    "dincae_adam_update_factored": (c_int, [vp, vp, vp, c_int, vp, c_size_t, c_int, vp, c_size_t, c_int,
                                            c_int, c_int, c_int, c_int, c_float, vp, c_float, c_float, c_float,
                                            vp, vp, vp]),
    "dincae_dense_factored_expand": (c_int, [vp, c_size_t, c_int, vp, c_size_t, c_int, c_int, c_int, c_int, c_int,
                                             vp, c_int, vp]),
    "dincae_dense_bias_grad": (c_int, [vp, c_int, c_int, c_int, vp, vp]),
    "dincae_adam_multi_workspace": (c_size_t, [c_size_t]),
    "dincae_adam_step_multi": (c_int, [vp, vp, c_size_t, c_int, vp, vp, c_size_t, c_float, vp, c_float, c_float,
                                       c_float, c_float, vp, vp, vp, vp, c_size_t, vp]),
    "dincae_sum_ranks_f64": (c_int, [vp, c_size_t, c_int, c_int, vp, vp]),
    "dincae_f64_to_hilo": (c_int, [vp, c_int, vp, vp]),
    "dincae_hilo_to_f64": (c_int, [vp, c_int, vp, vp]),
    "dincae_head_recon": (c_int, [vp, c_int, c_int, c_int, vp, vp, vp]),
    "dincae_head_recon_records": (c_int, [vp, c_int, c_int, c_int, vp, vp, c_float, c_int, c_int, vp, vp]),
    "dincae_head_loss_workspace": (c_size_t, [c_int, c_int, c_int]),
    "dincae_head_loss": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_float, vp, vp, vp, vp, vp,
                                 c_size_t, vp]),
    "dincae_adam_workspace": (c_size_t, [c_size_t]),
    "dincae_adam_step": (c_int, [vp, vp, vp, vp, c_size_t, c_size_t, c_float, c_float, vp, c_float,
                                 c_float, c_float, c_float, vp, vp, vp, c_size_t, vp]),
    "dincae_l2_half_sumsq": (c_int, [vp, c_size_t, vp, vp, c_size_t, vp]),
    "dincae_ensemble_accumulate": (c_int, [vp, vp, c_float, c_size_t, vp, vp, vp]),
    "dincae_ensemble_finalize": (c_int, [vp, vp, c_size_t, c_int, c_float, vp, vp, vp]),
}


class DincaeError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise DincaeError(
            "%s not found: build the CUDA extension first (python -m dincae_b200.build). "
            "dincae_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        lib = load()
        msg = lib.dincae_error_string(rc).decode()
        raise DincaeError("%s failed: %s (code %d, cuda error %d)" %
                          (what or "dincae call", msg, rc, lib.dincae_last_cuda_error()))


def ptr(t):
    """Device (or host) address of a torch tensor / None."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def doubles(values):
    arr = (c_double * len(values))(*[float(v) for v in values])
    return arr
