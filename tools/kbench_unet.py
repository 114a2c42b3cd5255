"""UNet-shaped GEMM / conv micro-benchmarks timed by CUDA-graph replay (so ~10 us kernels are not host-bound),
operands rotated through > L2-sized sets. Prints achieved TFLOP/s, effective GB/s and the cuBLAS / cuDNN time."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitron_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16


def graph_time(fns, reps=5):
    """fns: list of closures (one per rotated operand set); returns ms per call."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for f in fns:
                f()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(3):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (3 * reps * len(fns))


def gemm_case(M, N, K, glu=0, bias=True, residual=False, act=0):
    nrot = max(2, min(8, int(300e6 // ((M * K + M * N) * 2)) + 1))
    As = [torch.randn((M, K), device=dev).to(BF) for _ in range(nrot)]
    W = (torch.randn((N, K), device=dev) * 0.02).to(BF)
    b = torch.zeros((N,), device=dev, dtype=BF) if bias else None
    No = N // 2 if glu else N
    Rs = [torch.randn((M, No), device=dev).to(BF) for _ in range(nrot)] if residual else [None] * nrot
    outs = [torch.empty((M, No), device=dev, dtype=BF) for _ in range(nrot)]
    fns = [(lambda j=j: ops.gemm(As[j], W, bias=b, glu=glu, act=act, residual=Rs[j], out=outs[j])) for j in range(nrot)]
    ms = graph_time(fns)
    Wt = W.t()
    refs = [(lambda j=j: torch.matmul(As[j], Wt)) for j in range(nrot)]
    ms_ref = graph_time(refs)
    fl = 2.0 * M * N * K
    by = (M * K + N * K + M * No * (2 if residual else 1)) * 2
    return {"op": "gemm", "M": M, "N": N, "K": K, "glu": glu, "res": residual, "us": round(ms * 1e3, 1),
            "tflops": round(fl / ms / 1e9, 1), "gbs": round(by / ms / 1e6, 1), "cublas_us": round(ms_ref * 1e3, 1)}


def conv_case(nb, h, w, cin, cout, kh=3, kw=3):
    nrot = max(2, min(8, int(300e6 // (nb * h * w * (cin + cout) * 2)) + 1))
    xs = [torch.randn((nb, h, w, cin), device=dev).to(BF) for _ in range(nrot)]
    wt = ops.pack_conv_weight(torch.randn((cout, cin, kh, kw), device=dev) * 0.02)
    b = torch.zeros((cout,), device=dev, dtype=BF)
    fns = [(lambda j=j: ops.conv_nhwc(xs[j], wt, kh, kw, pad_h=kh // 2, pad_w=kw // 2, bias=b)) for j in range(nrot)]
    ms = graph_time(fns)
    fl = 2.0 * nb * h * w * cin * cout * kh * kw
    return {"op": "conv", "x": [nb, h, w, cin], "cout": cout, "k": [kh, kw], "us": round(ms * 1e3, 1),
            "tflops": round(fl / ms / 1e9, 1)}


def main():
    cases = [(40960, 320, 320, 0, False), (40960, 320, 320, 0, True), (40960, 2560, 320, 2, False), (40960, 2560, 320, 0, False),
             (40960, 1280, 320, 0, False), (40960, 320, 1280, 0, True), (40960, 960, 320, 0, False),
             (10240, 640, 640, 0, True), (10240, 5120, 640, 2, False), (10240, 5120, 640, 0, False), (10240, 640, 2560, 0, True),
             (2560, 1280, 1280, 0, True), (2560, 10240, 1280, 2, False), (2560, 10240, 1280, 0, False), (2560, 1280, 5120, 0, True)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        cases = cases[:4]
    for M, N, K, glu, res in cases:
        print(json.dumps(gemm_case(M, N, K, glu=glu, residual=res)), flush=True)
    for nb, h, w, ci, co, kh, kw in [(16, 40, 64, 320, 320, 3, 3), (16, 20, 32, 640, 640, 3, 3), (16, 10, 16, 1280, 1280, 3, 3),
                                     (16, 5, 8, 1280, 1280, 3, 3), (1, 16, 2560, 320, 320, 3, 1), (1, 16, 640, 640, 640, 3, 1),
                                     (1, 16, 160, 1280, 1280, 3, 1), (1, 16, 40, 1280, 1280, 3, 1)]:
        print(json.dumps(conv_case(nb, h, w, ci, co, kh, kw)), flush=True)


if __name__ == "__main__":
    with torch.no_grad():
        main()
