"""A/B of the GELU evaluation in the bf16 path: fitted sigmoid-quintic (default, 2 MUFU) vs hardware tanh (1 MUFU).
Prints the end-to-end errors against the oracle / golden vectors and the step time for both.
python tools/gelu_ab.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import models_ref as O  # noqa: E402
from tests.util import build_pair, native  # noqa: E402

lib = native().load()
ora, prod = build_pair("XLSR_AASIST", "bf16")
x = O.synth_waveforms(4, 64000, seed=2021)
with torch.no_grad():
    ref, rt = ora(x, taps=True) if False else (ora(x), None)
    feats_ref = ora.ssl_model.extract_feat(x)
for variant, name in ((0, "fitted sigmoid (default)"), (4, "hardware tanh")):
    lib.rtdf_debug_gelu_variant(variant)
    eng = prod.engine()
    eng.use_graph = False
    got, taps = eng.forward(x.cuda(), want_taps=True)
    d = float((got.cpu() - ref).abs().max())
    fe = (taps["feats"].cpu() - feats_ref)
    print(f"{name:28s} max|dlogit| {d:.3e}   feats: max|d| {float(fe.abs().max()):.4f} rms {float(fe.pow(2).mean().sqrt()):.5f}"
          f"  (ref rms {float(feats_ref.pow(2).mean().sqrt()):.3f})")
    xb = torch.randn(64, 64000, device="cuda") * 0.1
    for _ in range(3):
        eng.forward(xb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.forward(xb)
    e1.record()
    torch.cuda.synchronize()
    print(f"{'':28s} {e0.elapsed_time(e1) / 10:.3f} ms / step (B=64, eager launches)")
lib.rtdf_debug_gelu_variant(0)
