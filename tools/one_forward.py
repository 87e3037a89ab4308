"""Run a few scoring forwards (for ncu launch lists).  python tools/one_forward.py [B] [iters] [kind] [N] [layers]
kind: aasist | conformer; layers < 24 builds the truncated student (every kernel type appears, few launches)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "real-time-deepfake-speech-detection_b200"

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kind = sys.argv[3] if len(sys.argv) > 3 else "aasist"
N = int(sys.argv[4]) if len(sys.argv) > 4 else 64000
layers = int(sys.argv[5]) if len(sys.argv) > 5 else 24
torch.manual_seed(0)
if kind == "aasist":
    xa = importlib.import_module(PKG + ".models.xlsr_aasist")
    model = xa.XLSR_AASIST("cpu", None) if layers == 24 else xa.My_XLSR_AASIST("cpu", None, num_layers=layers)
else:
    cb = importlib.import_module(PKG + ".models.conformer_baseline")
    model = cb.Model("cpu", None) if layers == 24 else cb.MyModel("cpu", None, num_layers=layers, fixed_call=True)
model = model.cuda().eval()
native = importlib.import_module(PKG + ".rtdf_runtime.native")
x = torch.randn(B, N, device="cuda") * 0.1
eng = model.engine()
torch.cuda.synchronize()
n0 = native.load().rtdf_launch_count()
for i in range(iters):
    y = eng.forward(x)
torch.cuda.synchronize()
print("launches per forward:", (native.load().rtdf_launch_count() - n0) // iters, "logits[0]:", y[0].tolist())
