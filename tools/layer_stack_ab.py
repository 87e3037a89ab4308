"""p50 / p99 latency (ms) of single streaming calls that fit the layer-stack kernel (<= 64 frames in flight), pinned host
waveform -> host score, for the current RTDF_LAYER_STACK setting.  Run once with RTDF_LAYER_STACK=0 and once with =1:
python tools/layer_stack_ab.py [calls]"""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "real-time-deepfake-speech-detection_b200"
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 500
for kind, mod in (("XLSR_AASIST", ".models.xlsr_aasist"), ("Model", ".models.conformer_baseline")):
    torch.manual_seed(0)
    model = getattr(importlib.import_module(PKG + mod), kind)("cpu", None).cuda().eval()
    eng = model.engine()
    model.rtdf_frozen = True
    for B, n in ((1, 16000), (1, 8000), (1, 20800), (2, 8000)):
        g = torch.Generator().manual_seed(5)
        host_in = (0.1 * torch.randn(B, n, generator=g)).pin_memory()
        host_out = torch.empty(B).pin_memory()
        buf = eng.static_input(B, n)
        ts = []
        for i in range(100 + calls):
            t0 = time.perf_counter()
            buf.copy_(host_in, non_blocking=True)
            out = eng.forward_static(B, n)
            host_out.copy_(out[:, 1], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            if i >= 100:
                ts.append(1e3 * (time.perf_counter() - t0))
        ts.sort()
        print(f"RTDF_LAYER_STACK={os.environ.get('RTDF_LAYER_STACK', 'default')} {kind} B={B} N={n} ({B * eng.num_frames(n)} frames): "
              f"p50 {ts[len(ts) // 2]:.3f} p99 {ts[int(len(ts) * 0.99)]:.3f} ms  score {host_out.tolist()}", flush=True)
