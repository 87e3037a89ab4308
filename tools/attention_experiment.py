"""Where does the warp-specialised attention kernel spend its time?  Run with RTDF_ATTN_DEBUG = bit mask (see
AttWsParams::debug in csrc/attention.cu; results are wrong when any bit is set) and compare.  B = 64, T = 199, 16 heads.
python tools/attention_experiment.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import P, call, stream  # noqa: E402
from tools.gpu_bench_kernels import timeit  # noqa: E402

B, T = 64, 199
qkv = torch.randn(B * T, 3072, device="cuda").to(torch.bfloat16)
qkv[:, :1024] *= 0.25
ctx = torch.empty(B * T, 1024, dtype=torch.bfloat16, device="cuda")
fn = lambda: call("rtdf_attention", P(qkv), P(ctx), B, T, 16, 1, 0, stream())
cold = timeit(fn, iters=10)
warm = timeit(fn, iters=10, flush=False)
print(f"debug={os.environ.get('RTDF_ATTN_DEBUG', '0'):>3s}  cold {cold * 1e3:6.1f} us  warm {warm * 1e3:6.1f} us")
