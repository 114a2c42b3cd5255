#!/bin/bash
# round 2, call B: one-pass GroupNorm + v2 GEMM/conv (split-K via reduce kernel): parity, suite, microbench, warm breakdown, ncu
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 200 -p no:cacheprovider -k "gemm or conv or groupnorm" > gpurun_out/t_kernels.log 2>&1
echo "kernel tests exit=$?" | tee gpurun_out/summary_r2b.txt
tail -n 3 gpurun_out/t_kernels.log
timeout 500 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "all gpu tests exit=$?" | tee -a gpurun_out/summary_r2b.txt
tail -n 8 gpurun_out/t_all.log
timeout 200 python tools/kbench_unet.py > gpurun_out/kbench_unet_v2b.jsonl 2> gpurun_out/kbench_unet_v2b.err
timeout 120 python tools/kineto_unet.py v2b > gpurun_out/kineto_v2b.log 2>&1
head -24 gpurun_out/kineto_v2b.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_v2 -s 4 -c 4 -o gpurun_out/conv1280_v2 python tools/ncu_conv1280.py > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
