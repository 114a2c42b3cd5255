"""Tile-width sweep: times every UNet / LLM GEMM and conv shape with VB200_FORCE_BN = 256/160/128/64 (graph
replay) next to the automatic choice, to calibrate pick_block_n's efficiency table."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kbench_unet as K  # noqa: E402

GEMMS = [(40960, 320, 320, 0, True), (40960, 2560, 320, 2, False), (40960, 960, 320, 0, False), (40960, 320, 1280, 0, True),
         (10240, 640, 640, 0, True), (10240, 5120, 640, 2, False), (10240, 1920, 640, 0, False), (10240, 640, 2560, 0, True),
         (2560, 1280, 1280, 0, True), (2560, 10240, 1280, 2, False), (2560, 3840, 1280, 0, False), (2560, 1280, 5120, 0, True),
         (640, 1280, 1280, 0, True), (2056, 3072, 1024, 0, False), (2056, 1024, 4096, 0, True), (6144, 12288, 4096, 0, False),
         (6144, 4096, 4096, 0, True), (6144, 22016, 4096, 1, False), (6144, 4096, 11008, 0, True)]
CONVS = [(16, 40, 64, 320, 320, 3, 3), (16, 20, 32, 640, 640, 3, 3), (16, 10, 16, 1280, 1280, 3, 3), (16, 5, 8, 1280, 1280, 3, 3),
         (16, 10, 16, 2560, 1280, 3, 3), (16, 40, 64, 640, 320, 3, 3), (1, 16, 2560, 320, 320, 3, 1), (1, 16, 640, 640, 640, 3, 1),
         (1, 16, 160, 1280, 1280, 3, 1), (1, 16, 40, 1280, 1280, 3, 1)]
with torch.no_grad():
    for M, N, K_, glu, res in GEMMS:
        row = {"gemm": [M, N, K_], "glu": glu, "res": res}
        for bn in ("auto", 256, 160, 128, 64):
            if bn == "auto":
                os.environ.pop("VB200_FORCE_BN", None)
            else:
                os.environ["VB200_FORCE_BN"] = str(bn)
            row[str(bn)] = K.gemm_case(M, N, K_, glu=glu, residual=res)["us"]
        print(json.dumps(row), flush=True)
    for nb, h, w, ci, co, kh, kw in CONVS:
        row = {"conv": [nb, h, w, ci, co, kh, kw]}
        for bn in ("auto", 256, 160, 128, 64):
            if bn == "auto":
                os.environ.pop("VB200_FORCE_BN", None)
            else:
                os.environ["VB200_FORCE_BN"] = str(bn)
            row[str(bn)] = K.conv_case(nb, h, w, ci, co, kh, kw)["us"]
        print(json.dumps(row), flush=True)
os.environ.pop("VB200_FORCE_BN", None)
