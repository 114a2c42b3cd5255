"""GPU parity for SURVEY.md §8 rows a8 (UNetSD_I2VGen), a9 (DDIM/CFG) and a12 (GLIGEN gated
self-attention): CUDA path vs golden vectors from the unmodified reference and vs the CPU oracle at a
mid-size configuration. bf16 kernels vs fp32 reference: <= 5% of the reference inf-norm, <= 4% rel. L2."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def assert_close(got, ref, what, rel_inf=0.05, rel_l2=0.04):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = ref.abs().max().item() + 1e-6
    e_inf = (got - ref).abs().max().item() / scale
    e_l2 = ((got - ref).norm() / (ref.norm() + 1e-6)).item()
    assert e_inf < rel_inf and e_l2 < rel_l2, f"{what}: inf {e_inf:.4f} l2 {e_l2:.4f}"


def to_dev(d, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


@pytest.fixture(scope="module")
def unet_fx(cuda):
    from oracle.weights import seeded_state_dict
    from vitron_b200.unet_i2vgen import UNetSD_I2VGen
    fx = torch.load(os.path.join(GOLD, "unet_tiny.pt"), weights_only=False)
    sd = seeded_state_dict(fx["shapes"], fx["seed"], fx["gain"])
    m = UNetSD_I2VGen(**fx["cfg"], device=cuda)
    m.load_state_dict(sd)
    return fx, sd, m


def test_unet_forward_vs_reference_golden(cuda, unet_fx):
    fx, sd, m = unet_fx
    out = m(**to_dev(fx["inputs"], cuda))
    assert_close(out, fx["out"], "UNetSD_I2VGen tiny")
    # local_image given as [b, c, h, w] and b=1 slice
    i = fx["inputs"]
    out1 = m(i["x"][:1].to(cuda), i["t"][:1].to(cuda), y=i["y"][:1].to(cuda), image=i["image"][:1].to(cuda),
             local_image=i["local_image"][:1, :, 0].to(cuda), fps=i["fps"][:1].to(cuda))
    assert_close(out1, fx["out"][:1], "UNet b=1")


def test_ddim_sampler_math(cuda):
    """Sampler arithmetic in isolation: identical analytic 'model' on both sides, fp32 -> tight."""
    from oracle import restate_unet as U
    from vitron_b200.unet_i2vgen import DiffusionDDIM
    g = torch.Generator().manual_seed(0)
    noise = torch.randn((2, 4, 3, 5, 7), generator=g)
    bias = torch.randn((2, 4, 3, 5, 7), generator=g) * 0.1
    model = lambda xt, t, b=None: torch.tanh(xt) * 0.3 + b.to(xt.device) * (t.view(-1, 1, 1, 1, 1).float() / 1000.0).to(xt.device)
    ref = U.ddim_sample_loop(noise, model, [dict(b=bias), dict(b=-bias)], 9.0, 10)
    got = DiffusionDDIM().ddim_sample_loop(noise.to(cuda), model, [dict(b=bias), dict(b=-bias)], guide_scale=9.0, ddim_timesteps=10)
    assert torch.allclose(got.cpu(), ref, atol=1e-4, rtol=1e-4)


def test_ddim_with_unet_vs_oracle(cuda, unet_fx):
    """End to end: CUDA UNet inside the CUDA sampler vs oracle UNet inside the oracle sampler, mild
    guidance (a large CFG scale multiplies the bf16 error of y - u by ~2s)."""
    from oracle import restate_unet as U
    from vitron_b200.unet_i2vgen import DiffusionDDIM
    fx, sd, m = unet_fx
    d = fx["ddim"]
    omodel = lambda xt, t, **kw: U.unet_forward(sd, fx["cfg"], xt, t, **kw)
    ref = U.ddim_sample_loop(d["noise"], omodel, [d["cond"], d["uncond"]], 1.5, 4)
    got = DiffusionDDIM().ddim_sample_loop(d["noise"].to(cuda), m, [to_dev(d["cond"], cuda), to_dev(d["uncond"], cuda)],
                                           guide_scale=1.5, ddim_timesteps=4)
    assert_close(got, ref, "ddim 4 steps", rel_inf=0.08, rel_l2=0.06)


def test_unet_midsize_vs_oracle(cuda):
    """dim 128, mults (1,2,4): every block type incl. down/up sampling at 3 scales, f=8, 16x24 latent."""
    from oracle import restate_unet as U
    from oracle.weights import seeded_state_dict
    from vitron_b200.unet_i2vgen import UNetSD_I2VGen
    cfg = dict(in_dim=4, concat_dim=4, dim=128, y_dim=1024, context_dim=1024, out_dim=4, dim_mult=[1, 2, 4], num_heads=4,
               head_dim=64, num_res_blocks=1, attn_scales=[1.0, 0.5, 0.25], num_tokens=4)
    m = UNetSD_I2VGen(**cfg, device=cuda)
    # parameter shapes from the oracle's own plan (names = reference names)
    shapes = unet_shapes(cfg)
    sd = seeded_state_dict(shapes, 7, 0.4)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(3)
    rn = lambda *s: torch.randn(s, generator=g)
    b, f, h, w = 1, 8, 16, 24
    inp = dict(x=rn(b, 4, f, h, w), t=torch.tensor([321]), y=rn(b, 77, 1024), image=rn(b, 1, 1024),
               local_image=rn(b, 4, f, h, w), fps=torch.tensor([8]))
    ref = U.unet_forward(sd, cfg, **inp)
    out = m(**to_dev(inp, cuda))
    assert_close(out, ref, "UNet midsize")


def unet_shapes(cfg):
    from vitron_b200.param_shapes import unet_shapes as f
    return f(cfg)


def test_unet_shapes_helper_matches_reference_names():
    fx = torch.load(os.path.join(GOLD, "unet_tiny.pt"), weights_only=False)
    assert unet_shapes(fx["cfg"]) == {k: list(v) for k, v in fx["shapes"].items()}


test_unet_shapes_helper_matches_reference_names.pytestmark = []  # CPU test


def test_gligen_vs_reference_golden(cuda):
    from oracle.weights import seeded_state_dict
    from vitron_b200.gligen import BasicTransformerBlock
    fx = torch.load(os.path.join(GOLD, "gligen_tiny.pt"), weights_only=False)
    ctx, objs = fx["context"].to(cuda), fx["objs"].to(cuda)
    for c in fx["cases"]:
        sd = seeded_state_dict(c["shapes"], fx["seed"])
        blk = BasicTransformerBlock(c["C"], 768, 768, c["heads"], c["C"] // c["heads"], "gatedSA", device=cuda)
        blk.load_state_dict(sd)
        x = c["x"].to(cuda)
        assert_close(blk.fuser(x, objs), c["fuser_out"], f"GatedSelfAttentionDense C={c['C']}")
        assert_close(blk(x, ctx, objs), c["block_out"], f"BasicTransformerBlock C={c['C']}")


def test_gligen_full_size_vs_oracle(cuda):
    """N=4096 visual tokens + 30 grounding tokens, C=320, 8 heads x 40 (64x64 latent)."""
    from oracle import restate_gligen as G
    from oracle.weights import seeded_state_dict
    from vitron_b200.gligen import GatedSelfAttentionDense
    C, heads, N = 320, 8, 4096
    shapes = {"linear.weight": [C, 768], "linear.bias": [C], "norm1.weight": [C], "norm1.bias": [C], "norm2.weight": [C],
              "norm2.bias": [C], "alpha_attn": [], "alpha_dense": [], "attn.to_q.weight": [C, C], "attn.to_k.weight": [C, C],
              "attn.to_v.weight": [C, C], "attn.to_out.0.weight": [C, C], "attn.to_out.0.bias": [C],
              "ff.net.0.proj.weight": [8 * C, C], "ff.net.0.proj.bias": [8 * C], "ff.net.2.weight": [C, 4 * C], "ff.net.2.bias": [C]}
    sd = seeded_state_dict(shapes, 5)
    g = torch.Generator().manual_seed(2)
    x, objs = torch.randn((1, N, C), generator=g), torch.randn((1, 30, 768), generator=g)
    ref = G.gated_self_attention_dense(sd, "", x, objs, heads)
    m = GatedSelfAttentionDense(C, 768, heads, C // heads, device=cuda)
    m.load_state_dict(sd)
    assert_close(m(x.to(cuda), objs.to(cuda)), ref, "GatedSA 4096 tokens")
