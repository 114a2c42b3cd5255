#!/bin/bash
# GPU call 7: per-warp GEMM epilogue — whole suite with the mode on, then A/B on the small-K consumers and the UNet GEMM shapes.
mkdir -p gpurun_out
rm -f gpurun_out/summary7.txt
VB200_GEMM_EPILOGUE=1 timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_all7.log 2>&1
echo "all gpu tests (per-warp epilogue) exit=$?" | tee -a gpurun_out/summary7.txt
tail -n 8 gpurun_out/t_all7.log
for mode in 0 1; do
  VB200_GEMM_EPILOGUE=$mode timeout 200 python tools/bench_focal.py --no-seem > gpurun_out/bench_focal7_m$mode.jsonl 2> gpurun_out/bench_focal7_m$mode.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_focal7_m$mode.jsonl').readline()); print('epilogue $mode focalnet ms', d['focalnet_l']['ms_per_image'])" | tee -a gpurun_out/summary7.txt
  VB200_GEMM_EPILOGUE=$mode timeout 200 python tools/bench_gligen.py > gpurun_out/bench_gligen7_m$mode.jsonl 2> gpurun_out/bench_gligen7_m$mode.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_gligen7_m$mode.jsonl').readline())['gligen_unet_sd14']; print('epilogue $mode gligen ms', d['ms_graph_e2e'])" | tee -a gpurun_out/summary7.txt
  VB200_GEMM_EPILOGUE=$mode timeout 200 python tools/kbench_unet.py > gpurun_out/kbench_unet7_m$mode.jsonl 2> gpurun_out/kbench_unet7_m$mode.err
  echo "kbench_unet mode $mode exit=$?" | tee -a gpurun_out/summary7.txt
done
cat gpurun_out/summary7.txt
