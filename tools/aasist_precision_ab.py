"""A/B of the AASIST residual-encoder conv precision in bf16 mode: (hi,lo) bf16 operand pairs, 3 MMAs per product
(default) vs plain bf16.  Prints logit error vs the fp32 oracle, agreement of the GraphPool node selections and the
step time.  python tools/aasist_precision_ab.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import models_ref as O  # noqa: E402

PKG = "real-time-deepfake-speech-detection_b200"
rt = importlib.import_module(PKG + ".rtdf_runtime")

ora = O.build("XLSR_AASIST", seed=1024)
x = O.synth_waveforms(8, 64000, seed=2021)
with torch.no_grad():
    taps_ref = {}
    ref = ora(x, taps=taps_ref)
for impl, name in ((0, "hi/lo split (3 MMAs)"), (1, "plain bf16")):
    eng = rt.Engine(ora.state_dict(), "cuda", "aasist", 24, precision="bf16", aasist_conv_impl=impl, use_graph=False)
    got, taps = eng.forward(x.cuda(), want_taps=True)
    d = (got.cpu() - ref).abs()
    same_S = float((taps["idx_S"].cpu().long() == taps_ref["idx_S"]).float().mean()) if "idx_S" in taps_ref else float("nan")
    same_T = float((taps["idx_T"].cpu().long() == taps_ref["idx_T"]).float().mean()) if "idx_T" in taps_ref else float("nan")
    print(f"{name:22s} max|dlogit| {float(d.max()):.3e} mean {float(d.mean()):.3e}  idx_S agree {same_S:.3f} idx_T agree {same_T:.3f}")
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
    print(f"{'':22s} {e0.elapsed_time(e1) / 10:.3f} ms / step (B=64, eager launches)")
    eng.close()
