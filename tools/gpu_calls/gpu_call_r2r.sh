#!/bin/bash
# round 2, call R: SEEM head: folded position adds + inference mode without aux outputs (small-map attention masks)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_seem_gpu.py tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -q --timeout 300 -p no:cacheprovider -k "seem or resize or region" > gpurun_out/t_seem2.log 2>&1
echo "seem tests exit=$?" | tee gpurun_out/summary_r2r.txt
tail -n 12 gpurun_out/t_seem2.log
timeout 300 python tools/bench_cfg34.py --only cfg4 > gpurun_out/cfg4_r2r.jsonl 2> gpurun_out/cfg4_r2r.err
cat gpurun_out/cfg4_r2r.jsonl; tail -2 gpurun_out/cfg4_r2r.err
timeout 200 python tools/kineto_seem.py r_noaux noaux > gpurun_out/kineto_seem_r_noaux.log 2>&1
grep -v Warn gpurun_out/kineto_seem_r_noaux.log | head -30
