"""Launch the hottest kernel shapes a few times each (target for `ncu --set full -k regex:...`)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import P, call, stream  # noqa: E402

DEV = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
B, T = 64, 199
M = B * T
bf = torch.bfloat16
reps = 3
if which in ("all", "gemm"):
    for (N, K, act) in ((4096, 1024, 1), (3072, 1024, 0), (1024, 4096, 0)):
        A = torch.randn(M, K, device=DEV).to(bf)
        W = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(bf)
        bias = torch.randn(N, device=DEV)
        out = torch.empty(M, N, dtype=bf, device=DEV)
        for _ in range(reps):
            call("rtdf_gemm_bf16", P(A), P(W), M, N, K, P(bias), act, 1.0, None, None, P(out), 256, stream())
if which in ("all", "conv"):
    x = torch.randn(B, 12799, 512, device=DEV).to(bf)
    w = (torch.randn(512, 1536, device=DEV) / 40).to(bf)
    v1 = torch.randn(512, device=DEV)
    y = torch.empty(B, 6399, 512, dtype=bf, device=DEV)
    for _ in range(reps):
        call("rtdf_conv1d_ln_gelu_bf16", P(x), B, 12799, 3, 2, P(w), P(v1), P(v1), P(v1), 1e-5, P(y), 512, stream())
if which in ("all", "attn"):
    qkv = torch.randn(M, 3072, device=DEV).to(bf)
    ctx = torch.empty(M, 1024, dtype=bf, device=DEV)
    for _ in range(reps):
        call("rtdf_attention", P(qkv), P(ctx), B, T, 16, 1, 0, stream())
torch.cuda.synchronize()
print("done")
