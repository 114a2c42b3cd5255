#!/bin/bash
# One GPU call: parity tests of the widened rows (f1 FocalNet, f2 VAE, SEEM golden), their benches, a launch list
# of the FocalNet forward and the headline bench. Every step has its own timeout; logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 420 python -m pytest tests/test_zfocal_gpu.py tests/test_zvae_gpu.py tests/test_seem_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/t_widen.log 2>&1
echo "widen tests exit=$?" | tee -a gpurun_out/summary.txt
tail -n 25 gpurun_out/t_widen.log
timeout 240 python tools/bench_focal.py > gpurun_out/bench_focal.jsonl 2> gpurun_out/bench_focal.err
echo "bench_focal exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_focal.jsonl; tail -n 5 gpurun_out/bench_focal.err
timeout 240 python tools/bench_vae.py > gpurun_out/bench_vae.jsonl 2> gpurun_out/bench_vae.err
echo "bench_vae exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_vae.jsonl; tail -n 5 gpurun_out/bench_vae.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/focal_launches.csv python tools/bench_focal.py --profile > gpurun_out/ncu_focal.log 2>&1
echo "ncu focal exit=$?" | tee -a gpurun_out/summary.txt
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
cat gpurun_out/summary.txt
