"""One full-size UNetSD_I2VGen forward (f=16, 40x64 latent) for `ncu --metrics gpu__time_duration.sum`
launch lists: prints the number of vitron_b200 launches of the last forward so the tail of the CSV can be cut."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vitron_b200 import ops, param_shapes as PS  # noqa: E402
from vitron_b200.unet_i2vgen import UNetSD_I2VGen  # noqa: E402

dev = torch.device("cuda:0")
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
    l0 = ops.launch_count()
    print("MARK_BEGIN", flush=True)
    unet(x, t, **kw)
    torch.cuda.synchronize()
    print("vb launches in last forward:", ops.launch_count() - l0)
