#!/bin/bash
# round 2, call H: cluster-pair (A multicast) GEMM/conv variant: parity, A/B timing, UNet forward A/B; pending SEEM prompt + 7B tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 120 -p no:cacheprovider -k "cluster_pair" > gpurun_out/t_cluster.log 2>&1
echo "cluster tests exit=$?" | tee gpurun_out/summary_r2h.txt
tail -n 15 gpurun_out/t_cluster.log
timeout 400 python tools/kbench_gemm_modes.py > gpurun_out/kbench_modes_v2h.jsonl 2> gpurun_out/kbench_modes_v2h.err
cat gpurun_out/kbench_modes_v2h.jsonl; tail -3 gpurun_out/kbench_modes_v2h.err
for rb in 0 1 4 5; do
  timeout 200 python tools/kineto_unet.py rb$rb rb=$rb > gpurun_out/kineto_unet_rb$rb.log 2>&1
  head -12 gpurun_out/kineto_unet_rb$rb.log
done
timeout 900 python -m pytest tests/test_seem_gpu.py tests/test_fullsize_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/t_seem_full.log 2>&1
echo "seem+fullsize tests exit=$?" | tee -a gpurun_out/summary_r2h.txt
tail -n 15 gpurun_out/t_seem_full.log
