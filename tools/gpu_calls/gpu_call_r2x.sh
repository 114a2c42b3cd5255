#!/bin/bash
# round 2, call X: validation of the final tree: whole GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t_all_final.log 2>&1
echo "gpu tests exit=$?" | tee gpurun_out/summary_r2x.txt
tail -n 4 gpurun_out/t_all_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/summary_r2x.txt
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
echo "bench exit=$?" | tee -a gpurun_out/summary_r2x.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_final.json').read().strip().splitlines()[-1])
print('tokens/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['phases'], 'roofline', round(d['roofline']['frac'],3), d['roofline']['decode_step']['frac'])
u=d['unet']; print('unet', round(u['value'],2), 'e2e', round(u['e2e']['value'],2), 'gemm roofline', round(u['roofline']['frac'],3), 'traffic', u['roofline']['traffic'])
print('video', round(d['video_cfg2']['value'],2), 'cpu', d['cpu_baseline']['value'], u['cpu_baseline']['value'], d['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r02_final_ref.json 2>> gpurun_out/bench_r02_final.err
tail -c 600 gpurun_out/bench_r02_final_ref.json
