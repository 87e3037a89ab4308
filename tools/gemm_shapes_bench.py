"""Time the four projection GEMMs of a transformer layer at the timed batch through the C-ABI (CUDA events, operands rotated
over 4 buffer sets).  Used for A/B runs of kernel switches:  RTDF_TAIL_SLICES=0|1 python tools/gemm_shapes_bench.py [M]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import P, call, stream  # noqa: E402

DEV = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 199
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


tot = 0.0
for name, N, K, act, resid in (("qkv", 3072, 1024, 0, False), ("out", 1024, 1024, 0, True), ("fc1", 4096, 1024, 1, False),
                               ("fc2", 1024, 4096, 0, True)):
    A = [torch.randn(M, K, device=DEV).to(bf) for _ in range(NB)]
    W = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(bf)
    bias = torch.randn(N, device=DEV)
    if resid:
        x = [torch.randn(M, N, device=DEV) for _ in range(NB)]
        t = timed(lambda i: call("rtdf_gemm_bf16", P(A[i]), P(W), M, N, K, P(bias), act, 1.0, P(x[i]), P(x[i]), None, 2256, stream()))
    else:
        out = [torch.empty(M, N, dtype=bf, device=DEV) for _ in range(NB)]
        t = timed(lambda i: call("rtdf_gemm_bf16", P(A[i]), P(W), M, N, K, P(bias), act, 1.0, None, None, P(out[i]), 2256, stream()))
    tot += t
    print(f"{name:4s} N={N} K={K}: {t:7.1f} us  {2.0 * M * N * K / t / 1e6:6.0f} TF/s")
print(f"layer total {tot:7.1f} us   (RTDF_TAIL_SLICES={os.environ.get('RTDF_TAIL_SLICES', '1')}, M={M})")
