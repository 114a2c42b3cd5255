#!/bin/bash
# round 2, call N: A/B of the CTA-pair GEMM inside the real prefill (bench.py LLM leg), pair on (default) vs never (mode 9)
mkdir -p gpurun_out
for mode in 9 1 9 1; do
  VB200_GEMM_MODE=$mode timeout 400 python bench.py --no-unet --no-video --steps 3 > gpurun_out/bench_r2n_$mode.json 2> gpurun_out/bench_r2n_$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2n_$mode.json').read().strip().splitlines()[-1])
print('mode $mode', round(d['value'],1), {k: round(v,3) for k,v in d['phases'].items() if 'ms' in k})
PY
done
