"""Dependent-chain latency of small-M GEMMs (the SEEM mask decoder's / GLIGEN grounding tokens' regime): 24 links
x -> gemm (+ residual, + layernorm) -> x ... replayed from one CUDA graph (PDL on), time per link."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitron_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
LINKS = 24


def chain_time(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), ops.pdl(True):
        fn()
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(20):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 20 / LINKS * 1e3


with torch.no_grad():
    for M, N, K in [(101, 512, 512), (64, 512, 512), (60, 768, 768), (33, 1024, 1024), (17, 512, 512), (128, 320, 320)]:
        Ws = [(torch.randn((N, K), device=dev) * 0.03).to(BF) for _ in range(LINKS)]
        W2 = [(torch.randn((K, N), device=dev) * 0.03).to(BF) for _ in range(LINKS)] if N != K else None
        b = torch.zeros((N,), device=dev, dtype=BF)
        x0 = torch.randn((M, K), device=dev).to(BF)
        g_, be = torch.ones((N,), device=dev, dtype=BF), torch.zeros((N,), device=dev, dtype=BF)
        row = {"M": M, "N": N, "K": K}

        def plain():
            x = x0
            for i in range(LINKS):
                y = ops.gemm(x, Ws[i], bias=b)
                x = y if N == K else y[:, :K].contiguous() if N > K else x   # keep the dependency
            return x

        def with_ln(fused):
            x = x0
            for i in range(LINKS):
                y = ops.gemm(x, Ws[i], bias=b, residual=x)
                x = ops.layernorm(y, g_, be, 1e-5)
            return x
        if N == K:
            row["us_per_gemm"] = round(chain_time(plain), 2)
            row["us_per_gemm_plus_layernorm"] = round(chain_time(lambda: with_ln(False)), 2)
        print(json.dumps(row), flush=True)
