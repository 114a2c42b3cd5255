#!/bin/bash
# round 2, call C: fixed GN smem budget + generic-variant fp32 store; parity, suite, microbench, warm breakdown, bench line
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_kernels_gpu.py -q --timeout 200 -p no:cacheprovider -k "gemm or conv or groupnorm" > gpurun_out/t_kernels.log 2>&1
echo "kernel tests exit=$?" | tee gpurun_out/summary_r2c.txt
tail -n 3 gpurun_out/t_kernels.log
timeout 120 python tools/kineto_unet.py v2c > gpurun_out/kineto_v2c.log 2>&1
head -24 gpurun_out/kineto_v2c.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider --deselect tests/test_kernels_gpu.py > gpurun_out/t_all.log 2>&1
echo "all gpu tests (minus kernels file) exit=$?" | tee -a gpurun_out/summary_r2c.txt
tail -n 12 gpurun_out/t_all.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err
echo "bench exit=$?" | tee -a gpurun_out/summary_r2c.txt
tail -c 1500 gpurun_out/bench_r2c.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2c.json").read().strip().splitlines()[-1])
    u = d.get("unet", {})
    print("tokens/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "unet steps/s", round(u.get("value", 0), 2), "roof", u.get("roofline", {}).get("frac"))
except Exception as e:
    print("no bench line:", e)
PY
