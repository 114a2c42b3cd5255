"""Minimal driver for an `ncu --set full --import-source on` capture of the small-K UNet GEMMs (run under ncu with
-k regex:gemm_bf16 -s 4 -c 4): 40960x320x320 (bias), same + residual, 2560x1280x1280 + residual, GEGLU 40960x2560x320."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitron_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
cases = [(40960, 320, 320, 0, False), (40960, 320, 320, 0, True), (2560, 1280, 1280, 0, True), (40960, 2560, 320, 2, False)]
with torch.no_grad():
    for rep in range(2):
        for M, N, K, glu, res in cases:
            a = torch.randn((M, K), device=dev).to(BF)
            w = (torch.randn((N, K), device=dev) * 0.02).to(BF)
            b = torch.zeros((N,), device=dev, dtype=BF)
            r = torch.randn((M, N // 2 if glu else N), device=dev).to(BF) if res else None
            ops.gemm(a, w, bias=b, glu=glu, residual=r)
    torch.cuda.synchronize()
print("done")
