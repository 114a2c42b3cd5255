"""Shape-level time breakdown of one full-size UNetSD_I2VGen forward: every vitron_b200.ops call is timed with
CUDA events (synchronising per call, so the sum is kernel time without launch overlap) and aggregated by
(op, shapes). Also counts torch-side ops between them via the wall clock of the un-instrumented forward."""
import os, sys, json, collections, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vitron_b200 import ops, param_shapes as PS  # noqa: E402
from vitron_b200.unet_i2vgen import UNetSD_I2VGen  # noqa: E402
import vitron_b200.unet_i2vgen as U  # noqa: E402

dev = torch.device("cuda:0")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])  # count, ms, flops
NAMES = ["gemm", "conv_nhwc", "conv_nhwc_direct", "layernorm", "groupnorm_nhwc", "attention", "attention_short",
         "add_rowgroup", "upsample2x_nhwc", "add", "cfg_combine", "rmsnorm"]


def shape_key(name, args, kw):
    ts = [tuple(a.shape) for a in list(args) + list(kw.values()) if torch.is_tensor(a)]
    extra = {k: v for k, v in kw.items() if not torch.is_tensor(v) and k in ("glu", "act", "stride", "kh", "kw")}
    pos = [a for a in args if isinstance(a, int)]
    return f"{name} {ts[:3]} {pos} {extra}"


def flops(name, args, kw):
    if name == "gemm":
        a, w = args[0], args[1]
        return 2.0 * a.numel() // a.shape[-1] * a.shape[-1] * w.shape[0]
    if name == "conv_nhwc":
        x, wt = args[0], args[1]
        stride = kw.get("stride", 1)
        return 2.0 * x.shape[0] * (x.shape[1] // stride) * (x.shape[2] // stride) * wt.shape[0] * wt.shape[1]
    if name == "attention":
        q, k = args[0], args[1]
        return 4.0 * q.shape[0] * q.shape[2] * q.shape[1] * k.shape[1] * q.shape[3]
    return 0.0


def wrap(name):
    fn = getattr(ops, name)

    def w(*args, **kw):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        out = fn(*args, **kw)
        b.record()
        b.synchronize()
        e = agg[shape_key(name, args, kw)]
        e[0] += 1
        e[1] += a.elapsed_time(b)
        e[2] += flops(name, args, kw)
        return out
    return w


with torch.no_grad():
    unet = UNetSD_I2VGen(**bench.UNET_CFG, device=dev)
    unet.load_state_dict(PS.random_state_dict(PS.unet_shapes(bench.UNET_CFG), dev, seed=4))
    g = torch.Generator(device=dev).manual_seed(4)
    rn = lambda *s: torch.randn(s, generator=g, device=dev)
    x, local = rn(1, 4, 16, 40, 64), rn(1, 4, 16, 40, 64)
    kw = dict(y=rn(1, 77, 1024), image=rn(1, 1, 1024), local_image=local, fps=torch.tensor([16], device=dev))
    t = torch.tensor([981], device=dev)
    for _ in range(2):
        unet(x, t, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    unet(x, t, **kw)
    b.record()
    torch.cuda.synchronize()
    fwd_ms = a.elapsed_time(b)
    orig = {n: getattr(ops, n) for n in NAMES}
    for n in NAMES:
        setattr(ops, n, wrap(n))
    unet(x, t, **kw)
    torch.cuda.synchronize()
    for n in NAMES:
        setattr(ops, n, orig[n])

tot = sum(v[1] for v in agg.values())
print(json.dumps({"forward_ms_eager": round(fwd_ms, 2), "sum_of_op_ms": round(tot, 2), "n_ops": sum(v[0] for v in agg.values())}))
by_op = collections.defaultdict(float)
for k, v in agg.items():
    by_op[k.split(" ")[0]] += v[1]
print(json.dumps({k: round(v, 2) for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])}))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    tf = v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0
    print(f"{v[1]:8.3f} ms  n={v[0]:4d}  {v[1] / v[0] * 1000:8.1f} us  {tf:7.1f} TF  {k}")
