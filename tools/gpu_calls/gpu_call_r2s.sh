#!/bin/bash
# round 2, call S: low-latency small-M GEMM (+ fused LayerNorm across a cluster): parity, whole kernel suite, SEEM timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 60 -p no:cacheprovider -k "small_m or gemm_ln" > gpurun_out/t_smallm.log 2>&1
echo "small-M tests exit=$?" | tee gpurun_out/summary_r2s.txt
tail -n 25 gpurun_out/t_smallm.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_seem_gpu.py tests/test_fullsize_gpu.py tests/test_unet_gligen_gpu.py tests/test_zgligen_unet_gpu.py tests/test_vitron_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t_most.log 2>&1
echo "kernel+seem+fullsize+gligen+vitron tests exit=$?" | tee -a gpurun_out/summary_r2s.txt
tail -n 8 gpurun_out/t_most.log
timeout 300 python tools/bench_cfg34.py --only cfg4 > gpurun_out/cfg4_r2s.jsonl 2> gpurun_out/cfg4_r2s.err
cat gpurun_out/cfg4_r2s.jsonl; tail -2 gpurun_out/cfg4_r2s.err
timeout 200 python tools/kineto_seem.py s_noaux noaux > gpurun_out/kineto_seem_s_noaux.log 2>&1
grep -v Warn gpurun_out/kineto_seem_s_noaux.log | head -24
