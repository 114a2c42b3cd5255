"""Driver for an `ncu --set full` capture of the few-tile 1280-channel UNet convolutions (run under ncu with
-k regex:gemm_v2 -s 4 -c 4): 3x3 at 16x(10x16), Conv3d (3,1,1) at 16 frames x 160 px, both 100 tiles of 128x256."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitron_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
with torch.no_grad():
    for rep in range(2):
        for (nb, h, w, ci, co, kh, kw) in [(16, 10, 16, 1280, 1280, 3, 3), (1, 16, 160, 1280, 1280, 3, 1),
                                           (16, 20, 32, 640, 640, 3, 3), (16, 5, 8, 1280, 1280, 3, 3)]:
            x = torch.randn((nb, h, w, ci), device=dev).to(BF)
            wt = ops.pack_conv_weight(torch.randn((co, ci, kh, kw), device=dev) * 0.02)
            b = torch.zeros((co,), device=dev, dtype=BF)
            ops.conv_nhwc(x, wt, kh, kw, pad_h=kh // 2, pad_w=kw // 2, bias=b)
    torch.cuda.synchronize()
print("done")
