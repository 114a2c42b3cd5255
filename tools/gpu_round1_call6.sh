#!/bin/bash
# GPU call 6: full suite after the GEMM kernel rebuild (resident-B mode compiled in, off), the GEMM / conv / module
# tests with the mode forced on, and the FocalNet + GLIGEN benches with the mode off / automatic.
mkdir -p gpurun_out
rm -f gpurun_out/summary6.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_all6.log 2>&1
echo "all gpu tests (mode 0) exit=$?" | tee -a gpurun_out/summary6.txt
tail -n 6 gpurun_out/t_all6.log
VB200_GEMM_B_RESIDENT=2 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_zfocal_gpu.py tests/test_zvae_gpu.py tests/test_zgligen_unet_gpu.py tests/test_zclip_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -k "gemm or focalnet or vae or gligen or openclip or seem" > gpurun_out/t_res6.log 2>&1
echo "forced resident-B tests exit=$?" | tee -a gpurun_out/summary6.txt
tail -n 12 gpurun_out/t_res6.log
for mode in 0 1; do
  VB200_GEMM_B_RESIDENT=$mode timeout 200 python tools/bench_focal.py --no-seem > gpurun_out/bench_focal6_m$mode.jsonl 2> gpurun_out/bench_focal6_m$mode.err
  echo "bench_focal mode $mode exit=$?" | tee -a gpurun_out/summary6.txt
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_focal6_m$mode.jsonl').readline()); print('mode $mode focalnet ms', d['focalnet_l']['ms_per_image'])"
  VB200_GEMM_B_RESIDENT=$mode timeout 200 python tools/bench_gligen.py > gpurun_out/bench_gligen6_m$mode.jsonl 2> gpurun_out/bench_gligen6_m$mode.err
  cat gpurun_out/bench_gligen6_m$mode.jsonl
done
cat gpurun_out/summary6.txt
