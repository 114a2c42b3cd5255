#!/bin/bash
# GPU call 5: graph reuse across videos (pipeline + UNet tests), i2vgen e2e sample, ncu --set full of the small-K GEMMs.
mkdir -p gpurun_out
rm -f gpurun_out/summary5.txt
timeout 600 python -m pytest tests/test_zi2vgen_pipeline_gpu.py tests/test_unet_gligen_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_5.log 2>&1
echo "pipeline+unet tests exit=$?" | tee -a gpurun_out/summary5.txt
tail -n 15 gpurun_out/t_5.log
timeout 400 python tools/bench_i2vgen.py --samples 3 > gpurun_out/bench_i2vgen5.jsonl 2> gpurun_out/bench_i2vgen5.err
echo "bench_i2vgen exit=$?" | tee -a gpurun_out/summary5.txt
cat gpurun_out/bench_i2vgen5.jsonl; tail -n 8 gpurun_out/bench_i2vgen5.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 1 -c 7 -o gpurun_out/gemm_smallk_full -f python tools/bench_focal.py --profile > gpurun_out/ncu_gemm5.log 2>&1
echo "ncu gemm exit=$?" | tee -a gpurun_out/summary5.txt
cat gpurun_out/summary5.txt
