#!/bin/bash
# round 2, call Q: decode step with the one-wave split policy + 2-trip prefetch; SEEM head per-kernel breakdown under graph replay
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_vitron_gpu.py -q --timeout 300 -p no:cacheprovider -k "decode or generate or paged" > gpurun_out/t_decode2.log 2>&1
echo "decode tests exit=$?"; tail -2 gpurun_out/t_decode2.log
for i in 1 2; do
timeout 400 python bench.py --no-unet --no-video --steps 3 > gpurun_out/bench_r2q_$i.json 2> gpurun_out/bench_r2q_$i.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2q_$i.json').read().strip().splitlines()[-1])
print('run $i', round(d['value'],1), {k: round(v,3) for k,v in d['phases'].items() if 'ms' in k}, d['roofline']['decode_step'])
PY
done
timeout 200 python tools/kineto_seem.py q > gpurun_out/kineto_seem_q.log 2>&1
grep -v Warn gpurun_out/kineto_seem_q.log | head -45
