"""Does running two scoring forwards concurrently (two streams, two workspaces) raise throughput?  The back-end's
small kernels leave SMs idle; a second forward in flight can fill them.  python tools/two_stream_experiment.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "real-time-deepfake-speech-detection_b200"
xa = importlib.import_module(PKG + ".models.xlsr_aasist")
rt = importlib.import_module(PKG + ".rtdf_runtime")

torch.manual_seed(0)
model = xa.XLSR_AASIST("cpu", None).cuda().eval()
e1 = model.engine()
e2 = rt.Engine({k: v for k, v in model.state_dict().items()}, "cuda", "aasist", 24)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
xs = [torch.randn(B, 64000, device="cuda") * 0.1 for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for e, s in ((e1, s1), (e2, s2)):
    with torch.cuda.stream(s):
        for i in range(3):
            e.forward(xs[i])
torch.cuda.synchronize()


def run(n, two):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for i in range(n):
        e, s = ((e1, s1), (e2, s2))[i & 1] if two else (e1, s1)
        with torch.cuda.stream(s):
            e.forward(xs[i & 3])
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for two in (False, True, False, True):
    ms = run(20, two)
    print(f"{'two streams' if two else 'one stream '}: {ms:.3f} ms / forward  {B / ms * 1e3:.0f} utt/s")
