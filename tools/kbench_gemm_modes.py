"""A/B of the v2 GEMM modes on the UNet level-0 / level-1 shapes, L2-WARM (operands re-used: what the kernel sees inside
the UNet, where its input was just written) and L2-cold: default (non-resident B), resident B, and the main loop alone
(epilogue work skipped: wrong results, timing only)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitron_b200 import ops  # noqa: E402
from tools.kbench_unet import graph_time  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
CASES = [(6144, 12288, 4096, 0, False), (6144, 4096, 11008, 0, True), (40960, 320, 320, 0, False), (40960, 320, 320, 0, True), (40960, 960, 320, 0, False), (40960, 2560, 320, 2, False),
         (40960, 1280, 320, 0, False), (10240, 640, 640, 0, True), (10240, 1920, 640, 0, False), (10240, 5120, 640, 2, False),
         (40960, 320, 1280, 0, True), (2560, 1280, 1280, 0, True)]
with torch.no_grad():
    for M, N, K, glu, res in CASES:
        No = N // 2 if glu else N
        W = (torch.randn((N, K), device=dev) * 0.02).to(BF)
        b = torch.zeros((N,), device=dev, dtype=BF)
        row = {"M": M, "N": N, "K": K, "glu": glu, "res": res}
        for temp in ("warm", "cold"):
            nrot = 1 if temp == "warm" else max(2, min(8, int(300e6 // ((M * K + M * No) * 2)) + 1))
            As = [torch.randn((M, K), device=dev).to(BF) for _ in range(nrot)]
            Rs = [torch.randn((M, No), device=dev).to(BF) for _ in range(nrot)] if res else [None] * nrot
            outs = [torch.empty((M, No), device=dev, dtype=BF) for _ in range(nrot)]
            fns = [(lambda j=j: ops.gemm(As[j], W, bias=b, glu=glu, residual=Rs[j], out=outs[j])) for j in range(nrot)]
            for mode, (rb, dbg) in {"plain": (8, 0), "pair": (4, 0), "auto": (0, 0), "plain_noepi": (8, 1), "pair_noepi": (4, 1)}.items():
                ops.set_gemm_debug(rb, dbg)
                row[f"{temp}_{mode}_us"] = round(graph_time(fns) * 1e3, 1)
            ops.set_gemm_debug(1, 0)
        if not glu:
            Wt = W.t()
            A0 = torch.randn((M, K), device=dev).to(BF)
            row["warm_cublas_us"] = round(graph_time([lambda: torch.matmul(A0, Wt)]) * 1e3, 1)
        print(json.dumps(row), flush=True)
