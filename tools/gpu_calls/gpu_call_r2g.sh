#!/bin/bash
# round 2, call G: resident-B GEMM variant (parity + A/B timing incl. main loop alone), GN v3 microbench, per-op profile
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 90 -p no:cacheprovider -k "gemm or groupnorm or conv" > gpurun_out/t_kernels.log 2>&1
echo "kernel tests exit=$?" | tee gpurun_out/summary_r2g.txt
tail -n 4 gpurun_out/t_kernels.log
timeout 300 python tools/kbench_gemm_modes.py > gpurun_out/kbench_modes_v2g.jsonl 2> gpurun_out/kbench_modes_v2g.err
cat gpurun_out/kbench_modes_v2g.jsonl; tail -3 gpurun_out/kbench_modes_v2g.err
timeout 200 python tools/kbench_gn.py > gpurun_out/kbench_gn_v2g.jsonl 2> gpurun_out/kbench_gn_v2g.err
cat gpurun_out/kbench_gn_v2g.jsonl
timeout 200 python tools/kineto_unet_ops.py v2g > gpurun_out/kineto_ops_v2g.log 2>&1
head -30 gpurun_out/kineto_ops_v2g.log
