"""GroupNorm(32) micro-benchmark at the UNetSD_I2VGen shapes: CUDA-graph replay, L2-warm (one buffer re-used) and L2-cold
(operands rotated through > 300 MB). Prints us per launch and effective GB/s (read + write)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitron_b200 import ops  # noqa: E402
from tools.kbench_unet import graph_time  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
SHAPES = [(16, 2560, 320, 4), (1, 40960, 320, 4), (16, 2560, 320, 0), (16, 640, 640, 4), (1, 10240, 640, 4), (16, 160, 1280, 4),
          (1, 2560, 1280, 4), (16, 40, 1280, 4), (1, 640, 1280, 4), (16, 2560, 640, 4), (16, 2560, 960, 4), (16, 640, 1280, 4),
          (16, 640, 1920, 4), (16, 160, 2560, 4), (16, 40, 2560, 4)]
with torch.no_grad():
    for n, sp, c, act in SHAPES:
        w, b = torch.ones((c,), device=dev, dtype=BF), torch.zeros((c,), device=dev, dtype=BF)
        nbytes = n * sp * c * 2
        res = {"n": n, "spatial": sp, "c": c, "act": act, "mb": round(nbytes / 1e6, 1)}
        for mode in ("warm", "cold"):
            nrot = 1 if mode == "warm" else max(2, min(12, int(300e6 // (2 * nbytes)) + 1))
            xs = [torch.randn((n, sp, c), device=dev).to(BF) for _ in range(nrot)]
            outs = [torch.empty_like(x) for x in xs]
            fns = [(lambda j=j: ops.groupnorm_nhwc(xs[j], w, b, 32, 1e-5, act=act, out=outs[j])) for j in range(nrot)]
            ms = graph_time(fns)
            res[mode + "_us"] = round(ms * 1e3, 1)
            res[mode + "_gbs"] = round(2 * nbytes / ms / 1e6, 0)
        print(json.dumps(res), flush=True)
