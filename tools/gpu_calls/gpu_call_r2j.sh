#!/bin/bash
# round 2, call J: whole GPU suite, smoke, default bench line, UNet launch list with DRAM bytes, --set full of the UNet GEMMs, LLM launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t_all_r02.log 2>&1
echo "gpu tests exit=$?" | tee gpurun_out/summary_r2j.txt
tail -n 6 gpurun_out/t_all_r02.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee -a gpurun_out/summary_r2j.txt
timeout 900 python bench.py > gpurun_out/bench_r02_full.json 2> gpurun_out/bench_r02_full.err
echo "bench exit=$?" | tee -a gpurun_out/summary_r2j.txt
cat gpurun_out/bench_r02_full.json; tail -5 gpurun_out/bench_r02_full.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/unet_launches_r02.csv python tools/profile_unet.py > gpurun_out/profile_unet_r02.log 2>&1
echo "unet launch list exit=$?" | tee -a gpurun_out/summary_r2j.txt
tail -2 gpurun_out/profile_unet_r02.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_v2 -s 4 -c 4 -f -o gpurun_out/ncu_gemm_v2_r02 python tools/ncu_gemm_smallk.py > gpurun_out/ncu_gemm_v2_r02.log 2>&1
echo "ncu full exit=$?" | tee -a gpurun_out/summary_r2j.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/llm_launches_r02.csv python bench.py --profile --no-unet --no-video > gpurun_out/profile_llm_r02.log 2>&1
echo "llm launch list exit=$?" | tee -a gpurun_out/summary_r2j.txt
ls -la gpurun_out | tail -8
