"""ctypes binding of ``librtdf.so`` (the C-ABI declared in ``include/rtdf.h``).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, a
``RuntimeError`` carrying ``rtdf_last_error()`` is raised.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "librtdf.so")

c_void_p, c_int, c_float, c_size_t, c_longlong = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                                  ctypes.c_size_t, ctypes.c_longlong)

BACKEND_AASIST, BACKEND_CONFORMER, BACKEND_NONE = 0, 1, 2
PREC_BF16, PREC_FP32 = 0, 1
REGIME_AUTO, REGIME_THROUGHPUT = 0, 1
REGIMES = {"auto": REGIME_AUTO, "throughput": REGIME_THROUGHPUT}
ACT_NONE, ACT_GELU, ACT_SWISH, ACT_SELU = 0, 1, 2, 3


class ModelDesc(ctypes.Structure):
    _fields_ = [("backend", c_int), ("n_layers", c_int), ("precision", c_int), ("conf_emb", c_int),
                ("conf_heads", c_int), ("conf_kernel", c_int), ("conf_blocks", c_int), ("attention_impl", c_int),
                ("aasist_conv_impl", c_int), ("gat_impl", c_int)]


class GatWeights(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("att_w", "att_b", "a11", "a22", "a12", "with_t", "with_b", "without_t",
                                        "without_b", "bn_s", "bn_t")] + [("inv_temp", c_float)]


class Taps(ctypes.Structure):
    _fields_ = [("feats", c_void_p), ("hidden", c_void_p), ("idx_S", c_void_p), ("idx_T", c_void_p),
                ("layers", c_void_p)]


# name -> (restype, argtypes); every symbol declared in include/rtdf.h
SIGNATURES = {
    "rtdf_create": (c_int, [ctypes.POINTER(c_void_p), c_int, ctypes.POINTER(ModelDesc)]),
    "rtdf_load_weight": (c_int, [c_void_p, ctypes.c_char_p, c_void_p, ctypes.POINTER(ctypes.c_int64), c_int]),
    "rtdf_finalize": (c_int, [c_void_p]),
    "rtdf_set_regime": (c_int, [c_void_p, c_int]),
    "rtdf_destroy": (None, [c_void_p]),
    "rtdf_last_error": (ctypes.c_char_p, []),
    "rtdf_num_frames": (c_int, [c_int]),
    "rtdf_workspace_bytes": (c_int, [c_void_p, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "rtdf_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_size_t,
                             ctypes.POINTER(Taps), c_void_p]),
    "rtdf_frontend": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rtdf_backend": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, ctypes.POINTER(Taps), c_void_p]),
    "rtdf_preemph": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "rtdf_wave_layernorm": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "rtdf_conv0_ln_gelu": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p]),
    "rtdf_conv0_tc_scratch_bytes": (c_longlong, [c_int, c_int]),
    "rtdf_conv0_tc_ln_gelu": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                      c_void_p, c_void_p]),
    "rtdf_conv0_gn_workspace_floats": (c_longlong, [c_int, c_int]),
    "rtdf_conv0_gn_gelu": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "rtdf_layernorm_rows": (c_int, [c_void_p, c_int, c_longlong, c_int, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "rtdf_gemm_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "rtdf_gemm_bf16_rowln": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                     c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "rtdf_fold_ln_weight": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rtdf_gemm_bf16_xres": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "rtdf_cast_stats_rows": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_void_p, c_void_p]),
    "rtdf_gemm_bf16_lnfold": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_int,
                                      c_void_p, c_void_p, c_int, c_void_p]),
    "rtdf_gemm_plan_splits": (c_int, [c_int, c_int, c_int]),
    "rtdf_gemm_bf16_splitk": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rtdf_layernorm_accum_rows": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_void_p, c_float, c_void_p,
                                          c_void_p, c_void_p]),
    "rtdf_gemm_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "rtdf_conv1d_ln_gelu_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p]),
    "rtdf_posconv_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "rtdf_posconv_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rtdf_attention": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rtdf_conv_planes_tc": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                    ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                    c_int, c_void_p]),
    "rtdf_conformer_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rtdf_conformer_glu_dwconv": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_int, c_void_p]),
    "rtdf_launch_count": (c_longlong, []),
    "rtdf_debug_gelu_variant": (c_int, [c_int]),
    "rtdf_profile_begin": (c_int, []),
    "rtdf_profile_end": (c_int, [c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_int)]),
    "rtdf_fit_duration": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "rtdf_score_sink": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rtdf_roc_counts": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "rtdf_roc_crossing": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_longlong, c_void_p, c_void_p]),
    "rtdf_gat_rows": (c_int, [c_int, c_int, c_void_p, c_int, c_int, c_int, ctypes.POINTER(GatWeights), c_void_p, c_void_p,
                              ctypes.POINTER(GatWeights), c_void_p, c_int, c_void_p]),
    "rtdf_graph_pool": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load librtdf.so (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python {os.path.join(_PKG, 'build.py')}` "
            "(needs nvcc with sm_100a support). There is no CPU fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    msg = load().rtdf_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status, what):
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}): {last_error()}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())
