#!/bin/bash
# GPU call 2: the whole GPU suite (no -x: every failure is wanted), then the benches of the widened rows, the headline
# bench with the raw-image e2e arm, and an ncu --set full capture of the depthwise conv. Logs -> gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/summary2.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "all gpu tests exit=$?" | tee -a gpurun_out/summary2.txt
tail -n 30 gpurun_out/t_all.log
timeout 240 python tools/bench_focal.py > gpurun_out/bench_focal2.jsonl 2> gpurun_out/bench_focal2.err
echo "bench_focal exit=$?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/bench_focal2.jsonl; tail -n 5 gpurun_out/bench_focal2.err
timeout 240 python tools/bench_gligen.py > gpurun_out/bench_gligen.jsonl 2> gpurun_out/bench_gligen.err
echo "bench_gligen exit=$?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/bench_gligen.jsonl; tail -n 5 gpurun_out/bench_gligen.err
timeout 400 python bench.py > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench exit=$?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/bench2.json; tail -n 5 gpurun_out/bench2.err
timeout 240 ncu --set full --clock-control none --import-source on -k regex:dwconv_gelu_pair -s 8 -c 4 -o gpurun_out/dwconv_full -f python tools/bench_focal.py --profile > gpurun_out/ncu_dwconv.log 2>&1
echo "ncu dwconv exit=$?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/summary2.txt
