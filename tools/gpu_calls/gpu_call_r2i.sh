#!/bin/bash
# round 2, call I: where does the GEMM epilogue time go? (stores skipped vs all skipped) + raw store-pattern rates
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/store_pattern tools/micro/store_pattern.cu && timeout 120 /tmp/store_pattern > gpurun_out/store_pattern.jsonl
cat gpurun_out/store_pattern.jsonl
timeout 400 python tools/kbench_gemm_modes.py > gpurun_out/kbench_modes_v2i.jsonl 2> gpurun_out/kbench_modes_v2i.err
cat gpurun_out/kbench_modes_v2i.jsonl; tail -3 gpurun_out/kbench_modes_v2i.err
