"""Shared helpers for the tests: package import, C-ABI call wrappers, model construction."""
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "real-time-deepfake-speech-detection_b200"


def pkg(sub=""):
    return importlib.import_module(PKG_NAME + (("." + sub) if sub else ""))


def native():
    return pkg("rtdf_runtime.native")


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def P(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def call(name, *args):
    nat = native()
    lib = nat.load()
    nat.check(getattr(lib, name)(*args), name)


def gemm_bf16(A, W, bias=None, act=0, scale=1.0, resid=None, out="f32", variant=256, inplace=False):
    M, K = A.shape
    N = W.shape[0]
    if inplace:   # x += act(A W^T + b) * scale, updated in place (TMA reduce-add path)
        x = resid.clone()
        call("rtdf_gemm_bf16", P(A), P(W), M, N, K, P(bias), act, scale, P(x), P(x), None, variant, stream())
        return x
    # canary margins around the outputs: an out-of-bounds write of a tile epilogue / TMA store fails the check
    pad, canary = 4096, -7777.0
    raw32 = torch.full((M * N + 2 * pad,), canary, dtype=torch.float32, device=A.device) if out in ("f32", "both") else None
    raw16 = torch.full((M * N + 2 * pad,), canary, dtype=torch.bfloat16, device=A.device) if out in ("bf16", "both") else None
    o32 = raw32[pad:pad + M * N].view(M, N) if raw32 is not None else None
    o16 = raw16[pad:pad + M * N].view(M, N) if raw16 is not None else None
    call("rtdf_gemm_bf16", P(A), P(W), M, N, K, P(bias), act, scale, P(resid), P(o32), P(o16), variant, stream())
    for raw in (raw32, raw16):
        if raw is not None:
            assert bool((raw[:pad] == canary).all()) and bool((raw[-pad:] == canary).all()), \
                f"rtdf_gemm_bf16 variant {variant} ({M},{N},{K}) wrote outside its output"
    return o32 if out == "f32" else (o16 if out == "bf16" else (o32, o16))


def gemm_f32(A, W, bias=None, act=0, scale=1.0, resid=None):
    M, K = A.shape
    N = W.shape[0]
    o = torch.empty(M, N, dtype=torch.float32, device=A.device)
    call("rtdf_gemm_f32", P(A), P(W), M, N, K, P(bias), act, scale, P(resid), P(o), stream())
    return o


def build_pair(kind, precision, seed=1024, device="cuda", **kwargs):
    """(oracle model on CPU, product model on `device`) with identical seeded weights."""
    from oracle import models_ref as O
    ora = O.build(kind, seed=seed, **kwargs)
    if kind in ("XLSR_AASIST", "My_XLSR_AASIST"):
        cls = getattr(pkg("models.xlsr_aasist"), kind)
        prod = cls("cpu", None, **kwargs)
    else:
        mod = pkg("models.conformer_baseline")
        prod = mod.Model("cpu", None, **kwargs) if kind == "ConformerModel" else mod.MyModel("cpu", None, **kwargs)
    prod.load_state_dict(ora.state_dict(), strict=True)
    prod = prod.to(device).eval()
    prod.rtdf_precision = precision
    return ora, prod
