#!/bin/bash
# A/B the headline bench under different environment switches in ONE gpurun call (same box, same clocks):
#   tools/bench_ab.sh "RTDF_GEMM_2SM=1" "RTDF_GEMM_2SM=0" ...
for cfg in "$@"; do
  env $cfg timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', 'ms/step', round(d['ms_per_step'], 3), 'utt/s', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'gemm TF', round(d['roofline']['achieved'], 1))"
done
