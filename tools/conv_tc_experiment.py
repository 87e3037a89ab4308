"""Where does the time of the AASIST shifted-row conv go?  Times conv2 (64 -> 64, 6 taps, residual, fp32 + hi/lo outputs) at
the timed batch (64 x 44 x 68 plane rows) with parts of the kernel switched off (RTDF_CONVTC_DEBUG, wrong results).
    for m in 0 1 2 4 6 7 8 16 31; do RTDF_CONVTC_DEBUG=$m python tools/conv_tc_experiment.py; done"""
import ctypes
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import P, call, stream  # noqa: E402

DEV = "cuda"
B, Hp, W = 64, 44, 66
Wp = W + 2
rows = B * Hp * Wp
bf = torch.bfloat16
NB = 3
ci = co = 64
x_hi = [torch.randn(rows, ci, device=DEV).to(bf) for _ in range(NB)]
x_lo = [(torch.randn(rows, ci, device=DEV) * 1e-3).to(bf) for _ in range(NB)]
w_hi = (torch.randn(6, co, ci, device=DEV) / math.sqrt(6 * ci)).to(bf)
w_lo = (w_hi.float() * 1e-3).to(bf)
vec = torch.ones(co, device=DEV)
resid = [torch.randn(rows, co, device=DEV) for _ in range(NB)]
o32 = [torch.empty(rows, co, device=DEV) for _ in range(NB)]
oh = [torch.empty(rows, co, dtype=bf, device=DEV) for _ in range(NB)]
ol = [torch.empty(rows, co, dtype=bf, device=DEV) for _ in range(NB)]
shift = [(kh - 1) * Wp + kw - 1 for kh in range(2) for kw in range(3)]
sh = (ctypes.c_int * 6)(*shift)
sb = (ctypes.c_int * 6)(*([0] * 6))


def run(i):
    call("rtdf_conv_planes_tc", P(x_hi[i]), P(x_lo[i]), ci, rows, Hp, Wp, P(w_hi), P(w_lo), co, 6, sh, sb, 1, 42, P(vec), P(vec),
         P(vec), 3, P(resid[i]), P(vec), P(vec), 3, P(o32[i]), P(oh[i]), P(ol[i]), 3, stream())


for i in range(6):
    run(i % NB)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30):
    run(i % NB)
e1.record()
torch.cuda.synchronize()
print(f"RTDF_CONVTC_DEBUG={os.environ.get('RTDF_CONVTC_DEBUG', '0'):>3s}: conv2 64->64 at B=64: {1e3 * e0.elapsed_time(e1) / 30:7.1f} us")
