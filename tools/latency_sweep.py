"""p50 latency (ms) of single streaming calls for batch 1..8 x {1 s, 4 s} chunks in both kernel-selection regimes
(pinned host waveform -> Engine.static_input -> CUDA-graph replay -> host score).  python tools/latency_sweep.py [calls]"""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "real-time-deepfake-speech-detection_b200"
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 300
torch.manual_seed(0)
model = importlib.import_module(PKG + ".models.xlsr_aasist").XLSR_AASIST("cpu", None).cuda().eval()
eng = model.engine()
model.rtdf_frozen = True
for n in (16000, 64000):
    for B in (1, 2, 4, 8):
        row = []
        for regime in ("auto", "throughput"):
            host_in = (0.1 * torch.randn(B, n)).pin_memory()
            host_out = torch.empty(B).pin_memory()
            buf = eng.static_input(B, n, regime=regime)
            ts = []
            for i in range(50 + calls):
                t0 = time.perf_counter()
                buf.copy_(host_in, non_blocking=True)
                out = eng.forward_static(B, n, regime=regime)
                host_out.copy_(out[:, 1], non_blocking=True)
                torch.cuda.current_stream().synchronize()
                if i >= 50:
                    ts.append(1e3 * (time.perf_counter() - t0))
            ts.sort()
            row.append((ts[len(ts) // 2], ts[int(len(ts) * 0.99)]))
        print(f"B={B} N={n} ({B * eng.num_frames(n)} frames): auto p50 {row[0][0]:.3f} p99 {row[0][1]:.3f} ms | "
              f"throughput p50 {row[1][0]:.3f} p99 {row[1][1]:.3f} ms")
