#!/usr/bin/env bash
# compute-sanitizer pass over the per-kernel parity tests (run on a B200 box through gpurun, ONE tool per call --
# B200_PROFILING.md: several tools in one call have left the GPU unusable).
#
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck'     # or racecheck | synccheck | initcheck
#
# The selection keeps the run bounded: small shapes of every kernel family (the timed-batch tests are left out:
# under the sanitizer they would run for hours).  The log lands in gpurun_out/sanitize_<tool>.log; copy its summary
# to profiles/.
# NOTE (r02): compute-sanitizer is CLOSED on this GPU pool (gpurun answers exit 86, profiles/r02_sanitize_memcheck.log);
# out-of-bounds writes are caught by the canary margins around the outputs in tests/kernel_checks.py (Guarded) and
# tests/util.py (gemm_bf16) instead.
set -u
TOOL="${1:-memcheck}"
SEL="${2:-test_preemph or test_layernorm_rows or test_conv0_ln_gelu or test_gemm_bf16_tcgen05 or test_gemm_split_k or test_conv1d_implicit_gemm_ln_gelu or test_conv_planes_tcgen05 or test_posconv or test_attention or test_graph_pool or test_graph_attention_rows or test_conformer}"
mkdir -p gpurun_out
LOG="gpurun_out/sanitize_${TOOL}.log"
echo "# compute-sanitizer --tool ${TOOL}; pytest -k '${SEL} and not timed_batch'" > "${LOG}"
timeout 1400 compute-sanitizer --tool "${TOOL}" --error-exitcode 1 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "(${SEL}) and not timed_batch" >> "${LOG}" 2>&1
RC=$?
echo "# exit code ${RC}" >> "${LOG}"
grep -E "ERROR SUMMARY|passed|failed|error" "${LOG}" | tail -5
exit ${RC}
