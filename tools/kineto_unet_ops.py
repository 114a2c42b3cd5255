"""In-situ (warm, back-to-back, CUDA-graph replay) time of every vitron_b200 op of one full-size UNetSD_I2VGen forward, by
(op, shapes): the ordered list of ops.* calls recorded while the graph is captured is aligned with the ordered list of
vb:: kernels torch.profiler (CUPTI) reports for one replay. Unlike ncu's serialised cold-cache list these durations add up
to the measured forward. Usage: python tools/kineto_unet_ops.py [tag]"""
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vitron_b200 import ops, param_shapes as PS  # noqa: E402
from vitron_b200.unet_i2vgen import UNetSD_I2VGen  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "run"
dev = torch.device("cuda:0")
KERNELS = {"gemm": ("gemm_v2_kernel", "gemm_bf16_tcgen05", "gemv_bf16"), "conv_nhwc": ("gemm_v2_kernel", "gemm_bf16_tcgen05"),
           "groupnorm_nhwc": ("gn_onepass",), "layernorm": ("rownorm",), "attention": ("flash_attn",),
           "attention_short": ("attn_short",), "upsample2x_nhwc": ("upsample2x",), "cfg_combine": ("cfg_kernel",),
           "conv_nhwc_direct": ("conv_direct",), "add": ("add_kernel",)}
seq = []


def wrap(name):
    fn = getattr(ops, name)

    def w(*args, **kw):
        ts = [tuple(a.shape) for a in list(args) + list(kw.values()) if torch.is_tensor(a)]
        extra = {k: v for k, v in kw.items() if isinstance(v, (int, float, bool)) and k in ("glu", "act", "stride", "n")}
        flags = "+res" if kw.get("residual") is not None else ""
        pos = [a for a in args if isinstance(a, int)]
        seq.append((name, f"{name}{flags} {ts[:2]} {pos} {extra}", args, kw))
        return fn(*args, **kw)
    return w


def flops(name, args, kw):
    if name == "gemm":
        a, w = args[0], args[1]
        return 2.0 * (a.numel() // a.shape[-1]) * a.shape[-1] * w.shape[0]
    if name == "conv_nhwc":
        x, wt = args[0], args[1]
        st = kw.get("stride", 1)
        return 2.0 * x.shape[0] * (x.shape[1] // st) * (x.shape[2] // st if wt.shape[1] > 3 or args[3] > 1 else x.shape[2]) * wt.shape[0] * wt.shape[1] * x.shape[-1]
    if name == "attention":
        q, k = args[0], args[1]
        return 4.0 * q.shape[0] * q.shape[2] * q.shape[1] * k.shape[1] * q.shape[3]
    return 0.0


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
    orig = {n: getattr(ops, n) for n in KERNELS}
    for n in KERNELS:
        setattr(ops, n, wrap(n))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = unet(x, t, **kw)
    for n in KERNELS:
        setattr(ops, n, orig[n])
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        graph.replay()
        torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "vb::" in e.name]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
i = 0
unmatched = 0
for name, key, args, kw_ in seq:
    pats = KERNELS[name]
    if i >= len(evs) or not any(p in evs[i].name for p in pats):
        unmatched += 1
        continue
    dur = evs[i].device_time
    i += 1
    if i < len(evs) and "splitk_reduce" in evs[i].name and name in ("gemm", "conv_nhwc"):
        dur += evs[i].device_time
        i += 1
        key += " [split-K]"
    a = agg[key]
    a[0] += 1
    a[1] += dur
    a[2] += flops(name, args, kw_)
tot = sum(v[1] for v in agg.values())
print(json.dumps({"tag": tag, "ops": len(seq), "vb_kernels": len(evs), "aligned_kernels": i, "unmatched_ops": unmatched,
                  "sum_ms": round(tot / 1e3, 3)}))
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
for k, v in rows[:70]:
    tf = v[2] / (v[1] * 1e-6) / 1e12 if v[1] > 0 and v[2] > 0 else 0
    print(f"{v[1] / 1e3:7.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:3d} {v[1] / v[0]:7.1f} us {tf:7.1f} TF  {k}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump({k: v for k, v in rows}, open(f"gpurun_out/kineto_unet_ops_{tag}.json", "w"), indent=1)
