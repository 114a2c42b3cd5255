#!/bin/bash
# round 2, call U: small-M kernel v3 (64x64 tiles, whole-K prefetch, L2 weight prefetch): parity, chain latency, SEEM
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_vitron_gpu.py -q --timeout 200 -p no:cacheprovider -k "small_m or gemm_ln or midsize" > gpurun_out/t_smallm3.log 2>&1
echo "small-M tests exit=$?" | tee gpurun_out/summary_r2u.txt
tail -n 5 gpurun_out/t_smallm3.log
timeout 300 python tools/kbench_smallm.py > gpurun_out/kbench_smallm.jsonl 2> gpurun_out/kbench_smallm.err
cat gpurun_out/kbench_smallm.jsonl; tail -3 gpurun_out/kbench_smallm.err
timeout 300 python tools/bench_cfg34.py --only cfg4 > gpurun_out/cfg4_r2u.jsonl 2> gpurun_out/cfg4_r2u.err
cat gpurun_out/cfg4_r2u.jsonl; tail -2 gpurun_out/cfg4_r2u.err
VB200_GEMM_MODE=17 timeout 300 python tools/bench_cfg34.py --only cfg4 > gpurun_out/cfg4_r2u_nosmallm.jsonl 2>> gpurun_out/cfg4_r2u.err
cat gpurun_out/cfg4_r2u_nosmallm.jsonl
