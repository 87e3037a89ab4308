"""Time the dominant kernels in isolation at the BASELINE config-3 shapes (B=64, T=199) with CUDA
events on the launching stream; prints TFLOP/s or GB/s per kernel.  L2 is flushed between timed
iterations by writing a 256 MB buffer."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import P, call, stream  # noqa: E402

DEV = "cuda"
_flush = None


def timeit(fn, iters=8, warm=3, flush=True):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B, T = 64, 199
    M = B * T
    bf = torch.bfloat16
    print(f"device: {torch.cuda.get_device_name()}  M = {M}")
    for (N, K, act, name) in ((3072, 1024, 0, "qkv"), (1024, 1024, 0, "out_proj"), (4096, 1024, 1, "fc1+gelu"),
                              (1024, 4096, 0, "fc2"), (1024, 512, 0, "proj")):
        A = torch.randn(M, K, device=DEV).to(bf)
        W = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(bf)
        bias = torch.randn(N, device=DEV)
        out = torch.empty(M, N, dtype=bf, device=DEV)
        for v in (2256, 256, 128):
            ms = timeit(lambda: call("rtdf_gemm_bf16", P(A), P(W), M, N, K, P(bias), act, 1.0, None, None, P(out), v, stream()))
            print(f"gemm {name:9s} N={N:4d} K={K:4d} variant {v}: {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s")
        if act == 1:
            for gv, gname in ((4, "hw-tanh"), (5, "A&S erf")):
                ms = timeit(lambda: call("rtdf_gemm_bf16", P(A), P(W), M, N, K, P(bias), gv, 1.0, None, None, P(out), 256, stream()))
                print(f"gemm {name:9s} N={N:4d} K={K:4d} variant 256 gelu={gname}: {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s")
        ref = timeit(lambda: torch.nn.functional.linear(A, W))
        print(f"     torch/cuBLAS same shape (no epilogue):      {ref:7.3f} ms  {2.0 * M * N * K / ref / 1e9:8.1f} TFLOP/s")
    # conv feature encoder layers 1..6
    Ls = [12799, 6399, 3199, 1599, 799, 399, 199]
    ks = [3, 3, 3, 3, 2, 2]
    for i in range(6):
        Lin, Lout, k = Ls[i], Ls[i + 1], ks[i]
        x = torch.randn(B, Lin, 512, device=DEV).to(bf)
        w = (torch.randn(512, k * 512, device=DEV) / math.sqrt(512 * k)).to(bf)
        v1 = torch.randn(512, device=DEV)
        y = torch.empty(B, Lout, 512, dtype=bf, device=DEV)
        for v in (515, 516):
            ms = timeit(lambda: call("rtdf_conv1d_ln_gelu_bf16", P(x), B, Lin, k, 2, P(w), P(v1), P(v1), P(v1), 1e-5, P(y), v, stream()), iters=5)
            print(f"conv{i + 1} L_out={Lout:5d} k={k} variant {v}: {ms:7.3f} ms  {2.0 * B * Lout * 512 * 512 * k / ms / 1e9:8.1f} TFLOP/s")
        if i < 2:
            for gv, gname in ((4, "hw-tanh"), (5, "A&S erf")):
                call("rtdf_debug_gelu_variant", gv)
                ms = timeit(lambda: call("rtdf_conv1d_ln_gelu_bf16", P(x), B, Lin, k, 2, P(w), P(v1), P(v1), P(v1), 1e-5, P(y), 512, stream()), iters=5)
                print(f"conv{i + 1} L_out={Lout:5d} k={k} variant 512 gelu={gname}: {ms:7.3f} ms  {2.0 * B * Lout * 512 * 512 * k / ms / 1e9:8.1f} TFLOP/s")
            call("rtdf_debug_gelu_variant", 0)
    # conv0
    wav = torch.randn(B, 64000, device=DEV)
    wt = torch.randn(10, 512, device=DEV)
    v1 = torch.randn(512, device=DEV)
    y0 = torch.empty(B, Ls[0], 512, dtype=bf, device=DEV)
    ms = timeit(lambda: call("rtdf_conv0_ln_gelu", P(wav), B, 64000, P(wt), P(v1), P(v1), P(v1), 1e-5, None, P(y0), stream()))
    print(f"conv0+LN+GELU: {ms:7.3f} ms  write {y0.numel() * 2 / ms / 1e6:8.1f} GB/s")
    # layernorm 1024
    x = torch.randn(M, 1024, device=DEV)
    o = torch.empty(M, 1024, dtype=bf, device=DEV)
    g = torch.randn(1024, device=DEV)
    ms = timeit(lambda: call("rtdf_layernorm_rows", P(x), 0, M, 1024, P(g), P(g), 1e-5, 0, None, P(o), stream()))
    print(f"layernorm 1024 fp32->bf16: {ms * 1e3:7.1f} us  {M * 1024 * 6 / ms / 1e6:8.1f} GB/s")
    ms = timeit(lambda: call("rtdf_layernorm_rows", P(x), 0, M, 1024, P(g), P(g), 1e-5, 0, None, P(o), stream()), iters=20, flush=False)
    print(f"layernorm 1024 fp32->bf16 (input hot in L2): {ms * 1e3:7.1f} us  {M * 1024 * 6 / ms / 1e6:8.1f} GB/s")
    # attention
    qkv = torch.randn(M, 3072, device=DEV).to(bf)
    ctx = torch.empty(M, 1024, dtype=bf, device=DEV)
    for impl in (0, 2, 1):
        ms = timeit(lambda: call("rtdf_attention", P(qkv), P(ctx), B, T, 16, 1, impl, stream()), iters=5)
        print(f"attention impl {impl}: {ms:7.3f} ms  {4.0 * B * 16 * T * T * 64 / ms / 1e9:8.1f} TFLOP/s")
    # pos-conv
    xf = torch.randn(B, T, 1024, device=DEV)
    xb = xf.to(bf)
    w = (torch.randn(1024, 8192, device=DEV) / 90).to(bf)
    bias = torch.randn(1024, device=DEV)
    for impl in (0, 1):
        ms = timeit(lambda: call("rtdf_posconv_bf16", P(xf), P(xb), B, T, P(w), P(bias), impl, stream()), iters=5)
        print(f"posconv impl {impl}: {ms:7.3f} ms  {2.0 * M * 1024 * 64 * 128 / ms / 1e9:8.1f} TFLOP/s")


if __name__ == "__main__":
    main()
