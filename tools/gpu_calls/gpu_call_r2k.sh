#!/bin/bash
# round 2, call K: CTA-pair (cta_group::2) GEMM/conv: parity, A/B timing vs 1-CTA, UNet forward A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 60 -p no:cacheprovider -k "cluster_pair" > gpurun_out/t_pair.log 2>&1
echo "pair tests exit=$?" | tee gpurun_out/summary_r2k.txt
tail -n 25 gpurun_out/t_pair.log
if grep -q "passed" gpurun_out/t_pair.log && ! grep -q "failed" gpurun_out/t_pair.log; then
timeout 500 python tools/kbench_gemm_modes.py > gpurun_out/kbench_modes_v2k.jsonl 2> gpurun_out/kbench_modes_v2k.err
cat gpurun_out/kbench_modes_v2k.jsonl; tail -3 gpurun_out/kbench_modes_v2k.err
for rb in 8 4 0; do
  timeout 200 python tools/kineto_unet.py rb$rb rb=$rb > gpurun_out/kineto_unet_k_rb$rb.log 2>&1
  head -8 gpurun_out/kineto_unet_k_rb$rb.log | grep -v Warn
done
fi
