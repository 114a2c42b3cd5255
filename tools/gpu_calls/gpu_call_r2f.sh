#!/bin/bash
# round 2, call F: bias staged in smem (epilogue), split-KV tcgen05 attention, in-situ per-op breakdown, parity
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 90 -p no:cacheprovider > gpurun_out/t_kernels.log 2>&1
echo "kernel tests exit=$?" | tee gpurun_out/summary_r2f.txt
tail -n 4 gpurun_out/t_kernels.log
timeout 200 python tools/kineto_unet_ops.py v2f > gpurun_out/kineto_ops_v2f.log 2>&1
head -45 gpurun_out/kineto_ops_v2f.log
timeout 700 python -m pytest tests/test_fullsize_gpu.py tests/test_seem_gpu.py tests/test_unet_gligen_gpu.py tests/test_vitron_gpu.py tests/test_zfocal_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t_some.log 2>&1
echo "fullsize+seem+unet+vitron+focal tests exit=$?" | tee -a gpurun_out/summary_r2f.txt
tail -n 6 gpurun_out/t_some.log
timeout 200 python tools/bench_cfg34.py --only cfg4 > gpurun_out/cfg4_v2f.jsonl 2> gpurun_out/cfg4_v2f.err
cat gpurun_out/cfg4_v2f.jsonl
