"""Kernel micro-benchmarks (CUDA events, L2-cold via rotating operands): achieved TFLOP/s or GB/s
per kernel against MEASURED_PEAKS.json. Diagnostic; bench.py carries the contract line."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitron_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm_case(M, N, K, glu=0, nrot=4):
    As = [torch.randn((M, K), device=dev).to(BF) for _ in range(nrot)]
    Ws = [torch.randn((N, K), device=dev).to(BF) * 0.02 for _ in range(nrot)]
    i = [0]

    def f():
        j = i[0] % nrot
        i[0] += 1
        ops.gemm(As[j], Ws[j], glu=glu)
    ms = timed(f)
    def g():
        j = i[0] % nrot
        i[0] += 1
        torch.matmul(As[j], Ws[j].t())
    ms_ref = timed(g)
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    gb = (M * K + N * K + M * N) * 2 / (ms * 1e-3) / 1e9
    return {"op": "gemm", "M": M, "N": N, "K": K, "glu": glu, "ms": round(ms, 4), "tflops": round(tf, 1), "gbs": round(gb, 1),
            "cublas_ms": round(ms_ref, 4), "vs_cublas": round(ms_ref / ms, 3)}


def attn_case(B, H, S, D, causal):
    q, k, v = (torch.randn((B, S, H, D), device=dev).to(BF) for _ in range(3))
    ms = timed(lambda: ops.attention(q, k, v, causal=causal))
    fl = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
    qs, ks, vs = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    ms_ref = timed(lambda: torch.nn.functional.scaled_dot_product_attention(qs, ks, vs, is_causal=causal))
    return {"op": "attention", "B": B, "H": H, "S": S, "D": D, "causal": causal, "ms": round(ms, 4),
            "tflops": round(fl / (ms * 1e-3) / 1e12, 1), "sdpa_ms": round(ms_ref, 4)}


def conv_case(nb, h, w, cin, cout, k=3):
    x = torch.randn((nb, h, w, cin), device=dev).to(BF)
    wt = ops.pack_conv_weight(torch.randn((cout, cin, k, k), device=dev) * 0.02)
    ms = timed(lambda: ops.conv_nhwc(x, wt, k, k))
    fl = 2.0 * nb * h * w * cin * cout * k * k
    xc = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    wc = torch.randn((cout, cin, k, k), device=dev).to(BF).contiguous(memory_format=torch.channels_last)
    ms_ref = timed(lambda: torch.nn.functional.conv2d(xc, wc, padding=k // 2))
    return {"op": "conv", "nb": nb, "h": h, "w": w, "cin": cin, "cout": cout, "ms": round(ms, 4),
            "tflops": round(fl / (ms * 1e-3) / 1e12, 1), "cudnn_ms": round(ms_ref, 4)}


def main():
    res = []
    for M, N, K, glu in [(6144, 12288, 4096, 0), (6144, 4096, 4096, 0), (6144, 22016, 4096, 1), (6144, 4096, 11008, 0),
                         (8192, 8192, 8192, 0), (2056, 3072, 1024, 0), (2056, 4096, 1024, 0), (2056, 1024, 4096, 0),
                         (8, 12288, 4096, 0), (8, 4096, 4096, 0), (8, 22016, 4096, 1), (8, 4096, 11008, 0), (8, 32000, 4096, 0)]:
        res.append(gemm_case(M, N, K, glu))
        print(json.dumps(res[-1]), flush=True)
    for B, H, S, D, c in [(8, 32, 768, 128, True), (8, 16, 257, 64, False), (16, 5, 2560, 64, False)]:
        res.append(attn_case(B, H, S, D, c))
        print(json.dumps(res[-1]), flush=True)
    for nb, h, w, ci, co in [(16, 40, 64, 320, 320), (16, 20, 32, 640, 640), (16, 10, 16, 1280, 1280), (1, 256, 256, 512, 512)]:
        res.append(conv_case(nb, h, w, ci, co))
        print(json.dumps(res[-1]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kbench.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    with torch.no_grad():
        main()
