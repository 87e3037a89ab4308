"""Run every GPU parity check in its own subprocess (a CUDA fault in one check cannot poison the
others) with a per-check timeout, and print one PASS/FAIL line each plus the measured errors.

  python tools/gpu_diag.py [substring ...]      # run only checks whose name contains a substring
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECKS = [
    ("preemph", "tests.kernel_checks", "check_preemph", {}),
    ("wave_layernorm", "tests.kernel_checks", "check_wave_layernorm", {}),
    ("layernorm", "tests.kernel_checks", "check_layernorm", {}),
    ("conv0", "tests.kernel_checks", "check_conv0", {}),
    ("gemm_f32", "tests.kernel_checks", "check_gemm_f32", {}),
    ("gemm_bf16_v256", "tests.kernel_checks", "check_gemm_bf16", {"variants": (256,)}),
    ("gemm_bf16_v2256_pair", "tests.kernel_checks", "check_gemm_bf16", {"variants": (2256,)}),
    ("gemm_rowln", "tests.kernel_checks", "check_gemm_rowln", {}),
    ("gemm_bf16_v128", "tests.kernel_checks", "check_gemm_bf16", {"variants": (128,)}),
    ("gemm_bf16_v64", "tests.kernel_checks", "check_gemm_bf16", {"variants": (64,)}),
    ("gelu_epilogue", "tests.kernel_checks", "check_gelu_epilogue", {}),
    ("conv1d_tc_v512", "tests.kernel_checks", "check_conv1d_tc", {"variants": (512,)}),
    ("conv1d_tc_v513", "tests.kernel_checks", "check_conv1d_tc", {"variants": (513,)}),
    ("conv1d_tc_v514", "tests.kernel_checks", "check_conv1d_tc", {"variants": (514,)}),
    ("conv1d_tc_v515", "tests.kernel_checks", "check_conv1d_tc", {"variants": (515,)}),
    ("conv_planes_tc_x3", "tests.kernel_checks", "check_conv_planes_tc", {"nsplit": 3}),
    ("conv_planes_tc_x1", "tests.kernel_checks", "check_conv_planes_tc", {"nsplit": 1}),
    ("posconv", "tests.kernel_checks", "check_posconv", {}),
    ("attention_simt", "tests.kernel_checks", "check_attention", {"impls": (1,)}),
    ("attention_ws", "tests.kernel_checks", "check_attention", {"impls": (0,)}),
    ("attention_tile", "tests.kernel_checks", "check_attention", {"impls": (2,)}),
    ("graph_pool", "tests.kernel_checks", "check_graph_pool", {}),
    ("backend_block_fp32", "tests.e2e_checks", "check_backend_block", {"precision": "fp32"}),
    ("backend_block_bf16", "tests.e2e_checks", "check_backend_block", {"precision": "bf16"}),
    ("frontend_block_fp32", "tests.e2e_checks", "check_frontend_block", {"precision": "fp32"}),
    ("frontend_block_bf16", "tests.e2e_checks", "check_frontend_block", {"precision": "bf16"}),
    ("frontend_64000_fp32_L1", "tests.e2e_checks", "check_frontend_block", {"precision": "fp32", "N": 64000, "kind": "My_XLSR_AASIST", "num_layers": 1}),
    ("frontend_64000_bf16_L1", "tests.e2e_checks", "check_frontend_block", {"precision": "bf16", "N": 64000, "kind": "My_XLSR_AASIST", "num_layers": 1}),
    ("frontend_64000_bf16_L24", "tests.e2e_checks", "check_frontend_block", {"precision": "bf16", "N": 64000}),
    ("backend_block_fp32_64000", "tests.e2e_checks", "check_backend_block", {"precision": "fp32", "N": 64000, "B": 2}),
    ("e2e_aasist_fp32", "tests.e2e_checks", "check_e2e", {"precision": "fp32"}),
    ("e2e_aasist_bf16", "tests.e2e_checks", "check_e2e", {"precision": "bf16"}),
    ("e2e_conformer_fp32", "tests.e2e_checks", "check_e2e", {"kind": "ConformerModel", "precision": "fp32"}),
    ("e2e_conformer_bf16", "tests.e2e_checks", "check_e2e", {"kind": "ConformerModel", "precision": "bf16"}),
    ("golden_aasist_64000_bf16", "tests.e2e_checks", "check_golden", {"name": "xlsr_aasist_n64000_b2", "precision": "bf16"}),
    ("quirks_fp32", "tests.e2e_checks", "check_ragged_and_quirks", {"precision": "fp32"}),
]

RUNNER = """
import sys, json, importlib
sys.path.insert(0, {root!r})
import torch
mod = importlib.import_module({mod!r})
res = getattr(mod, {fn!r})(**{kw!r})
torch.cuda.synchronize()
print("RESULT " + json.dumps(res, default=str))
"""


def main():
    pats = sys.argv[1:]
    summary = []
    for name, mod, fn, kw in CHECKS:
        if pats and not any(p in name for p in pats):
            continue
        code = RUNNER.format(root=ROOT, mod=mod, fn=fn, kw=kw)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=420, cwd=ROOT)
            ok = r.returncode == 0
            tail = (r.stdout + r.stderr).strip().splitlines()
            info = [l for l in tail if l.startswith("RESULT ")]
            msg = info[-1][7:] if info else "\n".join(tail[-12:])
        except subprocess.TimeoutExpired:
            ok, msg = False, "TIMEOUT (420 s)"
        dt = time.time() - t0
        print(f"[{'PASS' if ok else 'FAIL'}] {name} ({dt:.1f}s) {msg[:3000]}", flush=True)
        summary.append((name, ok))
    bad = [n for n, ok in summary if not ok]
    print(f"SUMMARY: {len(summary) - len(bad)}/{len(summary)} passed; failed: {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
