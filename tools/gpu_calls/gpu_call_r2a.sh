#!/bin/bash
# round 2, call A: new GroupNorm + v2 GEMM/conv kernels: kernel parity, whole GPU suite, UNet microbench + warm breakdown
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 200 -p no:cacheprovider -k "gemm or conv or groupnorm" > gpurun_out/t_kernels.log 2>&1
echo "kernel tests exit=$?" | tee gpurun_out/summary_r2a.txt
tail -n 5 gpurun_out/t_kernels.log
timeout 500 python -m pytest tests -m gpu -q -x --timeout 200 -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "all gpu tests exit=$?" | tee -a gpurun_out/summary_r2a.txt
tail -n 5 gpurun_out/t_all.log
timeout 200 python tools/kbench_unet.py > gpurun_out/kbench_unet_v2.jsonl 2> gpurun_out/kbench_unet_v2.err
timeout 120 python tools/kineto_unet.py v2 > gpurun_out/kineto_v2.log 2>&1
head -20 gpurun_out/kineto_v2.log
