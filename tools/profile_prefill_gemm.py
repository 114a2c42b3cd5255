"""The four per-layer prefill GEMMs of Vicuna-7B at M = 8 x 768 and the UNet level-0 GEGLU GEMM, one launch each
after a warm-up, for `ncu --set full -k regex:gemm_bf16_tcgen05` (tensor-pipe utilisation evidence)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitron_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
d, f, M = 4096, 11008, 6144
x = torch.randn((M, d), device=dev).to(BF)
a = torch.randn((M, f), device=dev).to(BF)
w = dict(wqkv=(torch.randn((3 * d, d), device=dev) * 0.02).to(BF), wo=(torch.randn((d, d), device=dev) * 0.02).to(BF),
         wgu=(torch.randn((2 * f, d), device=dev) * 0.02).to(BF), wdown=(torch.randn((d, f), device=dev) * 0.02).to(BF))
xu = torch.randn((40960, 320), device=dev).to(BF)
wu = (torch.randn((2560, 320), device=dev) * 0.02).to(BF)
bu = torch.zeros((2560,), device=dev, dtype=BF)
with torch.no_grad():
    for _ in range(2):  # second round is the profiled one (ncu -s 5 -c 5)
        ops.gemm(x, w["wqkv"])
        ops.gemm(x, w["wo"], residual=x)
        ops.gemm(x, w["wgu"], glu=ops.GLU_SWIGLU)
        ops.gemm(a, w["wdown"], residual=x)
        ops.gemm(xu, wu, bias=bu, glu=ops.GLU_GEGLU)
torch.cuda.synchronize()
print("ok")
