#!/bin/bash
# Final check of the round: whole GPU suite on the reverted GEMM kernel + FocalNet timing.
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/t_final.log 2>&1
echo "final gpu tests exit=$?" | tee gpurun_out/summary_final.txt
tail -n 4 gpurun_out/t_final.log
timeout 40 python tools/bench_focal.py --no-seem > gpurun_out/bench_focal_final.jsonl 2> gpurun_out/bench_focal_final.err
python -c "import json; d=json.loads(open('gpurun_out/bench_focal_final.jsonl').readline()); print('focalnet ms', d['focalnet_l']['ms_per_image'])" | tee -a gpurun_out/summary_final.txt
