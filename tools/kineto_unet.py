"""Warm per-kernel time breakdown of one full-size UNetSD_I2VGen forward replayed from a CUDA graph, taken with
torch.profiler (CUPTI activity records: real back-to-back durations, warm L2 — unlike ncu's serialised cold-cache
list). Prints kernel-name totals and writes gpurun_out/kineto_unet.json. Usage: python tools/kineto_unet.py [tag]"""
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vitron_b200 import param_shapes as PS  # noqa: E402
from vitron_b200.unet_i2vgen import UNetSD_I2VGen  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "run"
dev = torch.device("cuda:0")
with torch.no_grad():
    unet = UNetSD_I2VGen(**bench.UNET_CFG, device=dev)
    unet.load_state_dict(PS.random_state_dict(PS.unet_shapes(bench.UNET_CFG), dev, seed=4))
    g = torch.Generator(device=dev).manual_seed(4)
    rn = lambda *s: torch.randn(s, generator=g, device=dev)
    nb = 2 if "b2" in sys.argv else 1   # b2: the batch-2 [cond | uncond] forward of the graphed CFG denoiser
    x, local = rn(nb, 4, 16, 40, 64), rn(nb, 4, 16, 40, 64)
    kw = dict(y=rn(nb, 77, 1024), image=rn(nb, 1, 1024), local_image=local, fps=torch.tensor([16] * nb, device=dev))
    t = torch.tensor([981] * nb, device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            unet(x, t, **kw)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    from vitron_b200 import ops
    for a_ in sys.argv:
        if a_.startswith("rb="):
            ops.set_gemm_debug(int(a_[3:]), 0)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph), ops.pdl("nopdl" not in sys.argv):
        out = unet(x, t, **kw)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(5):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    fwd_ms = a.elapsed_time(b) / 5
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        graph.replay()
        torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
for e in evs:
    k = e.name.split("(")[0][:90]
    agg[k][0] += 1
    agg[k][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
print(json.dumps({"tag": tag, "forward_ms_graph": round(fwd_ms, 3), "kernel_sum_ms": round(tot / 1e3, 3), "kernels": len(evs)}))
for k, v in rows[:30]:
    print(f"{v[1] / 1e3:8.3f} ms  {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  {v[1] / v[0]:7.1f} us  {k}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"tag": tag, "forward_ms_graph": fwd_ms, "kernel_sum_us": tot, "by_kernel": {k: v for k, v in rows}},
          open(f"gpurun_out/kineto_unet_{tag}.json", "w"), indent=1)
