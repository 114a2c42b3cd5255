"""Warm per-kernel time breakdown of the SEEM head (pixel decoder + mask decoder, 1024x1024 image, 101 queries) replayed from
its CUDA graph, taken with torch.profiler. Usage: python tools/kineto_seem.py [tag]"""
import collections, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitron_b200 import ops, param_shapes as PS  # noqa: E402
from vitron_b200.seem import MultiScaleMaskedTransformerDecoder, TransformerEncoderPixelDecoder, XDecoderHead  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "run"
dev = torch.device("cuda:0")
with torch.no_grad():
    in_ch = (192, 384, 768, 1536)
    sd = PS.random_state_dict(PS.seem_shapes(in_ch), dev, seed=3)
    head = XDecoderHead(TransformerEncoderPixelDecoder(in_ch, 512, 512, 8, 2048, 6, device=dev),
                        MultiScaleMaskedTransformerDecoder(512, 512, 101, 8, 2048, 9, 512, device=dev)).load_state_dict(sd)
    g = torch.Generator().manual_seed(3)
    feats = {f"res{i + 2}": torch.randn((1, c, 256 >> i, 256 >> i), generator=g).to(dev).to(torch.bfloat16) for i, c in enumerate(in_ch)}
    head.predictor.aux_outputs = "noaux" not in sys.argv
    head.enable_graph(True)
    for _ in range(3):
        head(feats)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(10):
        head(feats)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        head(feats)
        torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
seq = []
for e in evs:
    k = e.name.split("(")[0][:100]
    t = e.device_time if hasattr(e, "device_time") else e.cuda_time
    agg[k][0] += 1
    agg[k][1] += t
    seq.append((k, t))
tot = sum(v[1] for v in agg.values())
print(json.dumps({"tag": tag, "ms_graph": round(ms, 3), "kernel_sum_ms": round(tot / 1e3, 3), "kernels": len(evs)}))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1] / 1e3:8.3f} ms  {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  {v[1] / v[0]:7.1f} us  {k}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"tag": tag, "ms_graph": ms, "seq": seq}, open(f"gpurun_out/kineto_seem_{tag}.json", "w"))
