#!/bin/bash
# round 2, call W: LayerNorm folded into the consuming GEMM (UNet transformer blocks): parity, UNet tests, forward time A/B
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 120 -p no:cacheprovider -k "layernorm_fold or norms" > gpurun_out/t_lnfold.log 2>&1
echo "ln-fold tests exit=$?" | tee gpurun_out/summary_r2w.txt
tail -n 20 gpurun_out/t_lnfold.log
timeout 900 python -m pytest tests/test_unet_gligen_gpu.py tests/test_fullsize_gpu.py tests/test_zi2vgen_pipeline_gpu.py -q --timeout 600 -p no:cacheprovider -k "unet or i2vgen" > gpurun_out/t_unet_w.log 2>&1
echo "unet tests exit=$?" | tee -a gpurun_out/summary_r2w.txt
tail -n 6 gpurun_out/t_unet_w.log
timeout 200 python tools/kineto_unet.py w > gpurun_out/kineto_unet_w.log 2>&1
grep -v Warn gpurun_out/kineto_unet_w.log | head -12
timeout 200 python tools/kineto_unet_ops.py w > gpurun_out/kineto_ops_w.log 2>&1
grep -v Warn gpurun_out/kineto_ops_w.log | head -30
