#!/bin/bash
# round 2, call AD: GroupNorm sample chunking (slabs stay cached at batch 2): parity + UNet tests + steps/s
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gligen_gpu.py tests/test_zvae_gpu.py -q --timeout 300 -p no:cacheprovider -k "groupnorm or unet or vae" > gpurun_out/t_gn_ad.log 2>&1
echo "gn/unet/vae tests exit=$?" | tee gpurun_out/summary_r2ad.txt
tail -n 4 gpurun_out/t_gn_ad.log
timeout 600 python - <<'PY' 2>&1 | grep -v Warn | tail -4
import json, torch, bench
dev = torch.device("cuda:0")
with torch.no_grad():
    out = bench.bench_unet(dev, 1730.4, "measured", steps=10, with_cpu=False)
print(json.dumps({k: out[k] for k in ("value", "ms_per_step", "gpu_launches", "finite")}), out["e2e"]["value"], out["roofline"]["frac"])
PY
