#!/bin/bash
# round 2, call AB: software-pipelined decode attention (half-page double buffering): parity, standalone timing, decode step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_vitron_gpu.py tests/test_fullsize_gpu.py -q --timeout 400 -p no:cacheprovider -k "decode or generate or midsize or golden or vicuna" > gpurun_out/t_decode3.log 2>&1
echo "decode tests exit=$?" | tee gpurun_out/summary_r2ab.txt
tail -n 4 gpurun_out/t_decode3.log
timeout 200 python tools/profile_decode_attn.py 2>/dev/null | tee gpurun_out/decode_attn_r02_pipelined.json
for i in 1 2; do
timeout 400 python bench.py --no-unet --no-video --steps 3 > gpurun_out/bench_r2ab_$i.json 2> gpurun_out/bench_r2ab_$i.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2ab_$i.json').read().strip().splitlines()[-1])
print('run $i', round(d['value'],1), {k: round(v,3) for k,v in d['phases'].items() if 'ms' in k}, round(d['roofline']['decode_step']['frac'],3), d['tokens_check']['deterministic'])
PY
done
