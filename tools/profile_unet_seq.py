"""Run under `ncu --profile-from-start off --metrics gpu__time_duration.sum --csv`: one full-size UNet forward
inside cudaProfilerStart/Stop, with the ordered list of vitron_b200.ops calls (op, shapes) written to
gpurun_out/unet_seq.json so tools/join_unet_profile.py can attach shapes to the per-launch durations."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vitron_b200 import ops, param_shapes as PS  # noqa: E402
from vitron_b200.unet_i2vgen import UNetSD_I2VGen  # noqa: E402

dev = torch.device("cuda:0")
NAMES = ["gemm", "conv_nhwc", "conv_nhwc_direct", "layernorm", "groupnorm_nhwc", "attention", "attention_short",
         "add_rowgroup", "upsample2x_nhwc", "add", "cfg_combine", "rmsnorm"]
seq = []


def wrap(name):
    fn = getattr(ops, name)

    def w(*args, **kw):
        ts = [list(a.shape) for a in list(args) + list(kw.values()) if torch.is_tensor(a)]
        extra = {k: v for k, v in kw.items() if isinstance(v, (int, float, bool))}
        pos = [a for a in args if isinstance(a, (int, float))]
        seq.append({"op": name, "shapes": ts[:4], "pos": pos, "kw": extra})
        return fn(*args, **kw)
    return w


with torch.no_grad():
    unet = UNetSD_I2VGen(**bench.UNET_CFG, device=dev)
    unet.load_state_dict(PS.random_state_dict(PS.unet_shapes(bench.UNET_CFG), dev, seed=4))
    g = torch.Generator(device=dev).manual_seed(4)
    rn = lambda *s: torch.randn(s, generator=g, device=dev)
    x, local = rn(1, 4, 16, 40, 64), rn(1, 4, 16, 40, 64)
    kw = dict(y=rn(1, 77, 1024), image=rn(1, 1, 1024), local_image=local, fps=torch.tensor([16], device=dev))
    t = torch.tensor([981], device=dev)
    unet(x, t, **kw)
    torch.cuda.synchronize()
    for n in NAMES:
        setattr(ops, n, wrap(n))
    torch.cuda.profiler.start()
    unet(x, t, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(seq, open("gpurun_out/unet_seq.json", "w"))
print("ops recorded:", len(seq))
