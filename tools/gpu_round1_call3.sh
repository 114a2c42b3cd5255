#!/bin/bash
# GPU call 3: whole GPU suite after the epilogue / FocalNet kernel changes, FocalNet + headline benches, launch list.
mkdir -p gpurun_out
rm -f gpurun_out/summary3.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_all3.log 2>&1
echo "all gpu tests exit=$?" | tee -a gpurun_out/summary3.txt
tail -n 30 gpurun_out/t_all3.log
timeout 240 python tools/bench_focal.py > gpurun_out/bench_focal3.jsonl 2> gpurun_out/bench_focal3.err
echo "bench_focal exit=$?" | tee -a gpurun_out/summary3.txt
cat gpurun_out/bench_focal3.jsonl; tail -n 5 gpurun_out/bench_focal3.err
timeout 400 python bench.py > gpurun_out/bench3.json 2> gpurun_out/bench3.err
echo "bench exit=$?" | tee -a gpurun_out/summary3.txt
cat gpurun_out/bench3.json; tail -n 5 gpurun_out/bench3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/focal_launches3.csv python tools/bench_focal.py --profile > gpurun_out/ncu_focal3.log 2>&1
echo "ncu focal exit=$?" | tee -a gpurun_out/summary3.txt
timeout 200 python tools/bench_cfg34.py --only cfg3 > gpurun_out/bench_cfg3.jsonl 2> gpurun_out/bench_cfg3.err
echo "cfg3 exit=$?" | tee -a gpurun_out/summary3.txt
cat gpurun_out/bench_cfg3.jsonl
cat gpurun_out/summary3.txt
