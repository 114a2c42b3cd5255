#!/bin/bash
# round 2, call E: in-situ per-op UNet breakdown, GN microbench, new attention tests (mask / hd 40-80-160 on tcgen05), 7B test
mkdir -p gpurun_out
timeout 200 python tools/kineto_unet_ops.py v2e > gpurun_out/kineto_ops_v2e.log 2>&1
head -75 gpurun_out/kineto_ops_v2e.log
timeout 200 python tools/kbench_gn.py > gpurun_out/kbench_gn_v2e.jsonl 2> gpurun_out/kbench_gn_v2e.err
cat gpurun_out/kbench_gn_v2e.jsonl
timeout 400 python -m pytest tests/test_kernels_gpu.py -q --timeout 90 -p no:cacheprovider -k "attention or groupnorm" > gpurun_out/t_attn.log 2>&1
echo "attention/gn tests exit=$?" | tee gpurun_out/summary_r2e.txt
tail -n 4 gpurun_out/t_attn.log
timeout 400 python -m pytest tests/test_fullsize_gpu.py -q --timeout 300 -p no:cacheprovider -k vicuna > gpurun_out/t_7b.log 2>&1
echo "7B test exit=$?" | tee -a gpurun_out/summary_r2e.txt
tail -n 4 gpurun_out/t_7b.log
