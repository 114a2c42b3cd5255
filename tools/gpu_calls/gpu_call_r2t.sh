#!/bin/bash
# round 2, call T: whole-K-prefetch small-M kernel: parity + SEEM timing; bisect of the mid-size LLM test failure
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 60 -p no:cacheprovider -k "small_m or gemm_ln" > gpurun_out/t_smallm2.log 2>&1
echo "small-M tests exit=$?" | tee gpurun_out/summary_r2t.txt
tail -n 5 gpurun_out/t_smallm2.log
for cfg in "default" "VB200_GEMM_MODE=17" "VB200_DEC_SPLIT_KEYS=256" "VB200_GEMM_MODE=17 VB200_DEC_SPLIT_KEYS=256"; do
  if [ "$cfg" = "default" ]; then envs=""; else envs="$cfg"; fi
  env $envs timeout 300 python -m pytest tests/test_vitron_gpu.py -q --timeout 200 -p no:cacheprovider -k "midsize" > gpurun_out/t_mid.log 2>&1
  echo "midsize [$cfg] exit=$?" | tee -a gpurun_out/summary_r2t.txt
  grep "margin" gpurun_out/t_mid.log | tail -1 | cut -c1-200
done
timeout 300 python tools/bench_cfg34.py --only cfg4 > gpurun_out/cfg4_r2t.jsonl 2> gpurun_out/cfg4_r2t.err
cat gpurun_out/cfg4_r2t.jsonl; tail -2 gpurun_out/cfg4_r2t.err
VB200_GEMM_MODE=17 timeout 300 python tools/bench_cfg34.py --only cfg4 > gpurun_out/cfg4_r2t_nosmallm.jsonl 2>> gpurun_out/cfg4_r2t.err
cat gpurun_out/cfg4_r2t_nosmallm.jsonl
