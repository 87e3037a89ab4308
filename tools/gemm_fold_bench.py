"""Time the transformer GEMMs at the timed batch (M = 64 x 199 rows) with the plain epilogues against the folded-
LayerNorm epilogues (consumer: rstd (acc - mean c) + d; producer: x_old read + bf16 copy + row statistics), CUDA events,
operands rotated over 4 buffer sets (> L2).   python tools/gemm_fold_bench.py [variant]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import P, call, stream  # noqa: E402

DEV = "cuda"
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 2256
M = 64 * 199
bf = torch.bfloat16
NB = 4


def timed(fn, iters=40):
    for i in range(5):
        fn(i % NB)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % NB)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


for name, N, K, act in (("qkv", 3072, 1024, 0), ("fc1", 4096, 1024, 1)):
    A = [torch.randn(M, K, device=DEV).to(bf) for _ in range(NB)]
    W = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(bf)
    bias = torch.randn(N, device=DEV)
    cvec = torch.randn(N, device=DEV)
    stats = torch.zeros(M, 8, 2, device=DEV)
    stats[:, 0, 0] = 10.0
    stats[:, 0, 1] = 1200.0
    out = [torch.empty(M, N, dtype=bf, device=DEV) for _ in range(NB)]
    t0 = timed(lambda i: call("rtdf_gemm_bf16", P(A[i]), P(W), M, N, K, P(bias), act, 1.0, None, None, P(out[i]), variant, stream()))
    t1 = timed(lambda i: call("rtdf_gemm_bf16_lnfold", P(A[i]), P(W), M, N, K, P(cvec), P(bias), P(stats), 1e-5, act, None,
                              P(out[i]), variant, stream()))
    fl = 2.0 * M * N * K
    print(f"{name:5s} N={N} K={K}: plain {t0:7.1f} us ({fl / t0 / 1e6:6.0f} TF/s)   folded-LN consumer {t1:7.1f} us ({fl / t1 / 1e6:6.0f} TF/s)")
for name, N, K in (("out", 1024, 1024), ("fc2", 1024, 4096)):
    A = [torch.randn(M, K, device=DEV).to(bf) for _ in range(NB)]
    W = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(bf)
    bias = torch.randn(N, device=DEV)
    x = [torch.randn(M, N, device=DEV) for _ in range(NB)]
    xb = [torch.empty(M, N, dtype=bf, device=DEV) for _ in range(NB)]
    stats = torch.zeros(M, 8, 2, device=DEV)
    t0 = timed(lambda i: call("rtdf_gemm_bf16", P(A[i]), P(W), M, N, K, P(bias), 0, 1.0, P(x[i]), P(x[i]), None, variant, stream()))
    t1 = timed(lambda i: call("rtdf_gemm_bf16_xres", P(A[i]), P(W), M, N, K, P(bias), P(x[i]), P(xb[i]), P(stats), variant, stream()))
    ln = torch.empty(M, N, dtype=bf, device=DEV)
    g = torch.ones(N, device=DEV)
    t2 = timed(lambda i: call("rtdf_layernorm_rows", P(x[i]), 0, M, N, P(g), P(g), 1e-5, 0, None, P(ln), stream()))
    fl = 2.0 * M * N * K
    print(f"{name:5s} N={N} K={K}: in-place residual {t0:7.1f} us ({fl / t0 / 1e6:6.0f} TF/s)   + bf16 copy + stats {t1:7.1f} us "
          f"({fl / t1 / 1e6:6.0f} TF/s)   [stand-alone LayerNorm kernel {t2:5.1f} us]")
