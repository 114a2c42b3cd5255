"""GPU parity for SURVEY.md §8(f1): SEEM's FocalNet backbone (backbone/focal.py) on the vitron_b200 kernels.

Kernel level: every focal.cu entry point against a plain torch fp32 statement of the same op (tolerances per
test; bf16 storage, fp32 arithmetic). Module level: `vitron_b200.focal.FocalNet` against the golden outputs of
the UNMODIFIED reference class (tests/golden/focal_tiny.pt) and against the pinned CPU restatement
(oracle/restate_focal.py) on other configurations (pre-LN / post-LN, ragged image sizes); <= 4 % of the
reference inf-norm and <= 3 % relative L2 per output map. End to end: backbone -> pixel decoder -> mask decoder.
"""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

BF = torch.bfloat16
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rnd(shape, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).to(BF)


def close(a, b, atol, rtol, what=""):
    a, b = a.float().cpu(), b.float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs()
    bad = (err > atol + rtol * b.abs()).sum().item()
    assert bad == 0, f"{what}: {bad}/{a.numel()} mismatches, max err {err.max().item():.4g}"


def assert_close(got, ref, what, rel_inf=0.04, rel_l2=0.03):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = ref.abs().max().item() + 1e-6
    e_inf = (got - ref).abs().max().item() / scale
    e_l2 = ((got - ref).norm() / (ref.norm() + 1e-6)).item()
    assert e_inf < rel_inf and e_l2 < rel_l2, f"{what}: inf {e_inf:.4f} l2 {e_l2:.4f}"


# ------------------------------------------------------------------ kernels
@pytest.mark.parametrize("impl", [0, 1, 2, 3])
@pytest.mark.parametrize("k", [3, 5, 7, 9, 11])
@pytest.mark.parametrize("nb,h,w,c,ld_extra", [(1, 16, 24, 64, 0), (2, 9, 13, 72, 80), (1, 33, 5, 8, 8), (1, 64, 64, 192, 200),
                                               (1, 40, 300, 16, 0)])
def test_dwconv_gelu(cuda, k, nb, h, w, c, ld_extra, impl):
    from vitron_b200 import ops
    prev = ops.set_dwconv_impl(impl)
    try:
        _dwconv_case(cuda, k, nb, h, w, c, ld_extra)
    finally:
        ops.set_dwconv_impl(prev)


def _dwconv_case(cuda, k, nb, h, w, c, ld_extra):
    from vitron_b200 import ops
    ld = c + ld_extra
    buf = rnd((nb, h, w, ld), cuda, 1)
    off = 8 if ld_extra >= 16 else 0
    x = buf[..., off:off + c]
    wt = rnd((c, 1, k, k), cuda, 2, 1.0 / k)
    for act in (ops.ACT_GELU, ops.ACT_NONE):
        out = ops.dwconv_nhwc(x, ops.pack_dwconv_weight(wt), k, act=act)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), None, padding=k // 2, groups=c)
        if act == ops.ACT_GELU:
            ref = F.gelu(ref)
        close(out, ref.permute(0, 2, 3, 1), 2e-2, 1e-2, f"dwconv k{k} act{act}")


@pytest.mark.parametrize("nb,t,c", [(1, 384, 64), (2, 1000, 192), (1, 4096, 768), (3, 7, 1536), (1, 70000, 192)])
def test_colmean(cuda, nb, t, c):
    from vitron_b200 import ops
    x = rnd((nb, t, c), cuda, 3)
    x = (x.float() + 0.3).to(BF)
    for act in (ops.ACT_GELU, ops.ACT_NONE):
        out = ops.colmean(x.view(nb * t, c), nb, act=act)
        ref = x.float().mean(1)
        if act == ops.ACT_GELU:
            ref = F.gelu(ref)
        close(out, ref, 2e-4, 1e-3, "colmean")
    again = ops.colmean(x.view(nb * t, c), nb, act=ops.ACT_NONE)
    assert torch.equal(again, ops.colmean(x.view(nb * t, c), nb, act=ops.ACT_NONE)), "colmean must be deterministic"


@pytest.mark.parametrize("nb,t,c,L", [(1, 96, 64, 4), (2, 50, 192, 2), (1, 333, 1536, 1), (1, 17, 8, 6)])
def test_focal_modulate(cuda, nb, t, c, L):
    from vitron_b200 import ops
    levels = [rnd((nb * t, c), cuda, 10 + l) for l in range(L)]
    fo = rnd((nb * t, 2 * c + 8), cuda, 4)
    gates = fo[:, 2 * c:]
    glob = rnd((nb, c), cuda, 5).float().contiguous()
    out = ops.focal_modulate(levels, gates, glob, nb, 1.0 / (L + 1))
    ref = sum(levels[l].float() * gates[:, l:l + 1].float() for l in range(L))
    ref = ref + glob.repeat_interleave(t, 0) * gates[:, L:L + 1].float()
    close(out, ref / (L + 1), 1e-2, 1e-2, "focal_modulate")


def test_mul_rows_and_layernorm_add(cuda):
    from vitron_b200 import ops
    for rows, c in ((100, 64), (257, 192), (33, 1536), (5, 2048), (1000, 8)):
        fo = rnd((rows, 2 * c + 8), cuda, 6)
        b = rnd((rows, c), cuda, 7)
        close(ops.mul_rows(fo[:, :c], b), fo[:, :c].float() * b.float(), 1e-2, 1e-2, "mul_rows")
        x, res = rnd((rows, c), cuda, 8, 2.0), rnd((rows, c), cuda, 9)
        x = (x.float() + 1.5).to(BF)  # non-zero mean: exercises the centred variance
        w, bias = rnd((c,), cuda, 10), rnd((c,), cuda, 11)
        ref = res.float() + F.layer_norm(x.float(), (c,), w.float(), bias.float(), 1e-5)
        close(ops.layernorm_add(x, w, bias, res, 1e-5), ref, 3e-2, 1e-2, "layernorm_add")
        close(ops.layernorm_add(x, w, None, None, 1e-5), F.layer_norm(x.float(), (c,), w.float(), None, 1e-5), 3e-2, 1e-2,
              "layernorm (no residual)")
        y = res.clone()
        ops.layernorm_add(x, w, bias, y, 1e-5, out=y)  # in place on the residual stream
        close(y, ref, 3e-2, 1e-2, "layernorm_add in place")


@pytest.mark.parametrize("nb,c,h,w,k,stride,pad", [(1, 3, 64, 96, 7, 4, 2), (2, 3, 30, 45, 7, 4, 2), (1, 3, 28, 28, 4, 4, 0)])
def test_im2col_stem(cuda, nb, c, h, w, k, stride, pad):
    from vitron_b200 import ops
    g = torch.Generator().manual_seed(12)
    x = torch.randn((nb, c, h, w), generator=g).to(cuda)
    hp, wp = (h + 3) // 4 * 4, (w + 3) // 4 * 4
    ho, wo = (hp + 2 * pad - k) // stride + 1, (wp + 2 * pad - k) // stride + 1
    kpad = (c * k * k + 63) // 64 * 64
    for xin in (x, x.to(BF)):
        rows = ops.im2col_nchw(xin, k, stride, pad, ho, wo, kpad)
        xp = F.pad(xin.float(), (0, wp - w, 0, hp - h))
        ref = F.unfold(xp, k, padding=pad, stride=stride).transpose(1, 2).reshape(nb * ho * wo, c * k * k)
        assert torch.equal(rows[:, :c * k * k].float(), ref.to(BF).float()), "im2col must be exact (a gather)"
        assert rows[:, c * k * k:].abs().max().item() == 0


# ------------------------------------------------------------------ module
def build(cuda, cfg, sd):
    from vitron_b200.focal import FocalNet
    net = FocalNet(patch_size=cfg["patch_size"], embed_dim=cfg["embed_dim"], depths=cfg["depths"], mlp_ratio=cfg["mlp_ratio"],
                   patch_norm=cfg["patch_norm"], out_indices=cfg["out_indices"], focal_levels=cfg["focal_levels"],
                   focal_windows=cfg["focal_windows"], use_conv_embed=cfg["use_conv_embed"], use_postln=cfg["use_postln"],
                   use_postln_in_modulation=cfg["use_postln_in_modulation"], scaling_modulator=cfg["scaling_modulator"],
                   use_layerscale=cfg["use_layerscale"], device=cuda)
    return net.load_state_dict(sd)


def test_focalnet_vs_reference_golden(cuda):
    """Product path against outputs of the UNMODIFIED reference FocalNet (tests/golden/focal_tiny.pt)."""
    from oracle.weights import seeded_state_dict
    fx = torch.load(os.path.join(GOLD, "focal_tiny.pt"), weights_only=False)
    net = build(cuda, fx["cfg"], seeded_state_dict(fx["shapes"], fx["seed"]))
    outs = net(fx["x"].to(cuda))
    assert sorted(outs) == sorted(fx["outs"])
    for k, ref in fx["outs"].items():
        assert tuple(outs[k].shape) == tuple(ref.shape)
        assert_close(outs[k], ref, f"golden {k}")


@pytest.mark.parametrize("variant", ["postln_ragged", "preln", "focal_l_narrow"])
def test_focalnet_vs_oracle(cuda, variant):
    from oracle import restate_focal as FR
    from oracle.weights import seeded_state_dict
    from vitron_b200 import param_shapes
    base = dict(FR.FOCAL_L)
    if variant == "postln_ragged":      # image size not a multiple of 4 / odd maps at every stage
        cfg, size, nb = dict(base, embed_dim=64, depths=(1, 1, 2, 1)), (90, 62), 2
    elif variant == "preln":            # the other block ordering + LN inside the modulation, other windows
        cfg = dict(base, embed_dim=64, depths=(1, 1, 1, 1), use_postln=False, use_postln_in_modulation=True,
                   scaling_modulator=False, focal_levels=(3, 2, 2, 1), focal_windows=(5, 3, 7, 3))
        size, nb = (64, 64), 1
    else:                               # FocalNet-L widths, full block structure, shallow
        cfg, size, nb = dict(base, depths=(1, 1, 2, 1)), (128, 160), 1
    shapes = param_shapes.focalnet_shapes(cfg)
    sd = seeded_state_dict(shapes, 7)
    g = torch.Generator().manual_seed(2)
    x = torch.randn((nb, 3, *size), generator=g)
    ref = FR.focalnet_forward(sd, x, cfg)
    outs = build(cuda, cfg, sd)(x.to(cuda))
    for k, r in ref.items():
        assert_close(outs[k], r, f"{variant} {k}")


def test_seem_backbone_to_masks(cuda):
    """images -> FocalNet -> pixel decoder -> mask decoder against the oracle chain (tiny widths)."""
    from oracle import restate_focal as FR, restate_seem as S
    from oracle.weights import seeded_state_dict
    from vitron_b200 import param_shapes
    from vitron_b200.seem import MultiScaleMaskedTransformerDecoder, TransformerEncoderPixelDecoder, XDecoderHead
    cfg = dict(FR.FOCAL_L, embed_dim=64, depths=(1, 1, 2, 1))
    in_ch, C, ffn, Q, heads, dim_proj = (64, 128, 256, 512), 128, 256, 16, 2, 64
    bsd = seeded_state_dict(param_shapes.focalnet_shapes(cfg), 3)
    hsd = seeded_state_dict(S.seem_shapes(in_ch, C, ffn, Q, 1, 3, dim_proj), 4, 0.6)
    x = torch.randn((1, 3, 128, 160), generator=torch.Generator().manual_seed(6))
    feats_r = FR.focalnet_forward(bsd, x, cfg)
    mf_r, _, multi_r = S.pixel_decoder_forward(hsd, feats_r, "pixel_decoder.", nheads=heads, enc_layers=1)
    ref = S.mask_decoder_forward(hsd, multi_r, mf_r, "predictor.", heads=heads, num_layers=3)
    net = build(cuda, cfg, bsd)
    pd = TransformerEncoderPixelDecoder(in_ch, C, C, heads, ffn, 1, device=cuda)
    pr = MultiScaleMaskedTransformerDecoder(C, dim_proj, Q, heads, ffn, 3, C, device=cuda)
    head = XDecoderHead(pd, pr).load_state_dict(hsd)
    out = head(net(x.to(cuda)))
    assert_close(out["aux_outputs"][0]["pred_masks"], ref["aux_outputs"][0]["pred_masks"], "layer-0 masks", 0.08, 0.06)
    assert_close(out["pred_masks"], ref["pred_masks"], "pred_masks", 0.3, 0.1)
