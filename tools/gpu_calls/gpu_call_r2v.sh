#!/bin/bash
# round 2, call V: small-weight 17..64-row GEMMs on the persistent kernel: parity, chain latency, SEEM + GLIGEN timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 120 -p no:cacheprovider -k "gemm" > gpurun_out/t_gemm_v.log 2>&1
echo "gemm tests exit=$?" | tee gpurun_out/summary_r2v.txt
tail -n 4 gpurun_out/t_gemm_v.log
timeout 300 python tools/kbench_chain.py > gpurun_out/kbench_chain.jsonl 2> gpurun_out/kbench_chain.err
cat gpurun_out/kbench_chain.jsonl; tail -3 gpurun_out/kbench_chain.err
timeout 300 python tools/bench_cfg34.py --only cfg4 > gpurun_out/cfg4_r2v.jsonl 2> gpurun_out/cfg4_r2v.err
cat gpurun_out/cfg4_r2v.jsonl; tail -2 gpurun_out/cfg4_r2v.err
timeout 300 python tools/bench_gligen.py > gpurun_out/gligen_r2v.jsonl 2> gpurun_out/gligen_r2v.err
cat gpurun_out/gligen_r2v.jsonl; tail -2 gpurun_out/gligen_r2v.err
