#!/bin/bash
# GPU call 4: whole GPU suite (CLIP ViT-H width fix, i2vgen pipeline test, dwconv dispatch), FocalNet bench, i2vgen e2e sample.
mkdir -p gpurun_out
rm -f gpurun_out/summary4.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_all4.log 2>&1
echo "all gpu tests exit=$?" | tee -a gpurun_out/summary4.txt
tail -n 30 gpurun_out/t_all4.log
timeout 240 python tools/bench_focal.py > gpurun_out/bench_focal4.jsonl 2> gpurun_out/bench_focal4.err
echo "bench_focal exit=$?" | tee -a gpurun_out/summary4.txt
cat gpurun_out/bench_focal4.jsonl; tail -n 5 gpurun_out/bench_focal4.err
timeout 400 python tools/bench_i2vgen.py > gpurun_out/bench_i2vgen.jsonl 2> gpurun_out/bench_i2vgen.err
echo "bench_i2vgen exit=$?" | tee -a gpurun_out/summary4.txt
cat gpurun_out/bench_i2vgen.jsonl; tail -n 8 gpurun_out/bench_i2vgen.err
cat gpurun_out/summary4.txt
