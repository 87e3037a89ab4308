"""Where does the CTA-pair GEMM spend its time?  Run with RTDF_GEMM_DEBUG = 0 | 1 (no TMA loads) | 2 (no MMAs) |
4 (no epilogue) | combinations; prints the time of the four per-layer shapes (B=64, T=199), L2 flushed per iteration
and back to back (warm L2).  python tools/gemm_experiment.py"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import P, call, stream  # noqa: E402
from tools.gpu_bench_kernels import timeit  # noqa: E402

M = 64 * 199
bf = torch.bfloat16
dbg = os.environ.get("RTDF_GEMM_DEBUG", "0")
for (N, K, act, name) in ((3072, 1024, 0, "qkv"), (1024, 1024, 0, "out_proj"), (4096, 1024, 1, "fc1+gelu"), (1024, 4096, 0, "fc2")):
    A = torch.randn(M, K, device="cuda").to(bf)
    W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(bf)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, dtype=bf, device="cuda")
    x = torch.zeros(M, N, device="cuda")
    if name in ("out_proj", "fc2"):
        fn = lambda: call("rtdf_gemm_bf16", P(A), P(W), M, N, K, P(bias), act, 1.0, P(x), P(x), None, 2256, stream())
    else:
        fn = lambda: call("rtdf_gemm_bf16", P(A), P(W), M, N, K, P(bias), act, 1.0, None, None, P(out), 2256, stream())
    cold = timeit(fn, iters=8)
    warm = timeit(fn, iters=8, flush=False)
    print(f"debug={dbg} {name:9s} cold {cold * 1e3:7.1f} us  warm {warm * 1e3:7.1f} us  ({2.0 * M * N * K / warm / 1e9:7.1f} TFLOP/s warm)")
