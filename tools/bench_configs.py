"""Secondary BASELINE.json configurations on one B200 (the headline config lives in bench.py):
  C2  distilled students (6 XLS-R layers) at batch 256, 4 s utterances: My_XLSR_AASIST and MyModel (Conformer)
  C5  streaming chunks: 1 s / 4 s at batch 1, 2, 4, 8 for Model ("Conformer*") and XLSR_AASIST ("AASIST-SSL*"):
      p50 / p99 latency of model(x) measured with CUDA events and with the host clock (call + sync)
Prints one JSON object per line.  python tools/bench_configs.py [c2] [c5] [--calls N]"""
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "real-time-deepfake-speech-detection_b200"
xa = importlib.import_module(PKG + ".models.xlsr_aasist")
cb = importlib.import_module(PKG + ".models.conformer_baseline")


def build(kind):
    torch.manual_seed(1024)
    if kind == "XLSR_AASIST":
        m = xa.XLSR_AASIST("cpu", None)
    elif kind == "Conformer":
        m = cb.Model("cpu", None)
    elif kind == "Student6_AASIST":
        m = xa.My_XLSR_AASIST("cpu", None, num_layers=6, order="first")
    elif kind == "Student6_Conformer":
        m = cb.MyModel("cpu", None, num_layers=6, fixed_call=True)
    else:
        raise ValueError(kind)
    m = m.cuda().eval()
    m.rtdf_precision = "bf16"
    m.engine()
    m.rtdf_frozen = True
    return m


def throughput(kind, B, N, steps=10, warm=3):
    m = build(kind)
    xs = [torch.randn(B, N, device="cuda") * 0.1 for _ in range(3)]
    with torch.no_grad():
        for i in range(warm):
            m(xs[i % 3])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            m(xs[i % 3])
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return {"config": "C2", "model": kind, "batch": B, "n_samples": N, "ms_per_step": ms, "utt_per_s": B / ms * 1e3,
            "dtype": "bf16"}


def latency(kind, B, N, calls, warm=200):
    m = build(kind)
    x = torch.randn(B, N, device="cuda") * 0.1
    ev, wall = [], []
    with torch.no_grad():
        for _ in range(warm):
            m(x)
        torch.cuda.synchronize()
        for _ in range(calls):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            y = m(x)
            b.record()
            b.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            ev.append(a.elapsed_time(b))
    ev.sort()
    wall.sort()
    pick = lambda v, q: v[min(len(v) - 1, int(q * len(v)))]
    return {"config": "C5", "model": kind, "batch": B, "n_samples": N, "calls": calls,
            "p50_ms_cuda_events": pick(ev, 0.5), "p99_ms_cuda_events": pick(ev, 0.99),
            "p50_ms_host_clock": pick(wall, 0.5), "p99_ms_host_clock": pick(wall, 0.99),
            "chunks_per_s_at_p50": B / pick(wall, 0.5) * 1e3, "dtype": "bf16"}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    calls = 2000
    if "--calls" in sys.argv:
        calls = int(sys.argv[sys.argv.index("--calls") + 1])
    which = args or ["c2", "c5"]
    if "c2" in which:
        for kind in ("Student6_AASIST", "Student6_Conformer"):
            print(json.dumps(throughput(kind, 256, 64000)), flush=True)
    if "c5" in which:
        for kind in ("XLSR_AASIST", "Conformer"):
            for N in (16000, 64000):
                for B in (1, 2, 4, 8):
                    print(json.dumps(latency(kind, B, N, calls)), flush=True)


if __name__ == "__main__":
    main()
