#!/bin/bash
# round 2, call P: decode attention keys-per-split A/B (128 / 256 / 512) inside the real decode step
mkdir -p gpurun_out
for keys in 256 128 512 256 512; do
  VB200_DEC_SPLIT_KEYS=$keys timeout 400 python bench.py --no-unet --no-video --steps 3 > gpurun_out/bench_r2p_$keys.json 2> gpurun_out/bench_r2p_$keys.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2p_$keys.json').read().strip().splitlines()[-1])
print('keys $keys', round(d['value'],1), {k: round(v,3) for k,v in d['phases'].items() if 'ms' in k}, d['tokens_check']['deterministic'])
PY
done
