#!/bin/bash
# round 2, call Y: classifier-free guidance as ONE batch-2 UNet forward [cond | uncond]: parity tests + UNet steps/s
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gligen_gpu.py tests/test_zi2vgen_pipeline_gpu.py tests/test_fullsize_gpu.py -q --timeout 600 -p no:cacheprovider -k "unet or i2vgen or ddim" > gpurun_out/t_unet_y.log 2>&1
echo "unet tests exit=$?" | tee gpurun_out/summary_r2y.txt
tail -n 6 gpurun_out/t_unet_y.log
timeout 600 python - <<'PY' 2>&1 | grep -v Warn | tail -8
import json, torch, bench
dev = torch.device("cuda:0")
with torch.no_grad():
    out = bench.bench_unet(dev, 1730.4, "measured", steps=10, with_cpu=False)
print(json.dumps({k: out[k] for k in ("value", "ms_per_step", "gpu_launches", "achieved_tflops_per_gpu", "finite")}), json.dumps(out["e2e"]))
PY
