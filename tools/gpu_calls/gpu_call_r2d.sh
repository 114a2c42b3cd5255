#!/bin/bash
# round 2, call D: GN bulk-copy load + whole-warp blocks, residual prefetch ring, PDL everywhere on the UNet path, SEEM graph
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 90 -p no:cacheprovider -k "gemm or conv or groupnorm or norms or attention" > gpurun_out/t_kernels.log 2>&1
echo "kernel tests exit=$?" | tee gpurun_out/summary_r2d.txt
tail -n 4 gpurun_out/t_kernels.log
timeout 120 python tools/kineto_unet.py v2d > gpurun_out/kineto_v2d.log 2>&1
head -22 gpurun_out/kineto_v2d.log
timeout 120 python tools/kineto_unet.py v2d_nopdl nopdl > gpurun_out/kineto_v2d_nopdl.log 2>&1
head -4 gpurun_out/kineto_v2d_nopdl.log | tail -2
timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_seem_gpu.py tests/test_unet_gligen_gpu.py tests/test_vitron_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t_some.log 2>&1
echo "fullsize+seem+unet+vitron tests exit=$?" | tee -a gpurun_out/summary_r2d.txt
tail -n 6 gpurun_out/t_some.log
timeout 200 python tools/kbench_unet.py > gpurun_out/kbench_unet_v2d.jsonl 2> gpurun_out/kbench_unet_v2d.err
timeout 200 python tools/bench_cfg34.py --only cfg4 > gpurun_out/cfg4_v2d.jsonl 2> gpurun_out/cfg4_v2d.err
cat gpurun_out/cfg4_v2d.jsonl
