#!/bin/bash
# Runs the kernel numerics tests group by group, each in its own process with a hard timeout, so a
# hung kernel in one group cannot hide the results of the others. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in gemm_plain gemm_swap gemm_bias gemm_glu rowbias conv norms groupnorm "attention and not short" attention_short rope splice vision region; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 400 python -m pytest tests/test_kernels_gpu.py -m gpu -k "$grp" -q --timeout 150 --timeout-method thread -p no:cacheprovider > "gpurun_out/k_${name}.log" 2>&1
  echo "$name exit=$?" >> gpurun_out/k_summary.txt
  tail -n 3 "gpurun_out/k_${name}.log"
done
cat gpurun_out/k_summary.txt
