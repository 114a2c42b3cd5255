#!/bin/bash
# round 2, call M: decode attention with L2 page prefetch: parity + decode step time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_vitron_gpu.py -q --timeout 300 -p no:cacheprovider -k "decode or generate or paged or llama or vitron" > gpurun_out/t_decode.log 2>&1
echo "decode tests exit=$?" | tee gpurun_out/summary_r2m.txt
tail -n 5 gpurun_out/t_decode.log
timeout 600 python bench.py --no-unet --no-video > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2m.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['phases'], d['roofline']['frac'], d['roofline']['decode_step'])
PY
