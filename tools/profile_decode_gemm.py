"""Runs the four per-layer decode GEMMs (M=8) of Vicuna-7B over distinct weights so that
`ncu --set full -k regex:gemm_bf16_tcgen05 -s 8 -c 4` captures steady-state launches of the dominant
kernel (gemm_bf16_tcgen05_kernel<16,10>)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitron_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
d, f = 4096, 11008
x = torch.randn((8, d), device=dev).to(torch.bfloat16)
a = torch.randn((8, f), device=dev).to(torch.bfloat16)
layers = []
for _ in range(4):
    layers.append(dict(wqkv=torch.randn((3 * d, d), device=dev).to(torch.bfloat16), wo=torch.randn((d, d), device=dev).to(torch.bfloat16),
                       wgu=torch.randn((2 * f, d), device=dev).to(torch.bfloat16), wdown=torch.randn((d, f), device=dev).to(torch.bfloat16)))
with torch.no_grad():
    for L in layers:
        ops.gemm(x, L["wqkv"])
        ops.gemm(x, L["wo"])
        ops.gemm(x, L["wgu"], glu=ops.GLU_SWIGLU)
        ops.gemm(a, L["wdown"])
torch.cuda.synchronize()
print("ok")
