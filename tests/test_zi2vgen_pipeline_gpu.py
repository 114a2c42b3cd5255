"""GPU: the i2vgen-xl sampling chain (inference_i2vgen_entrance.py:118-209) through vitron_b200.i2vgen_pipeline —
OpenCLIP embedders -> VAE encode -> CFG DDIM loop over UNetSD_I2VGen (CUDA-graph replay) -> VAE decode — against the
same chain built from the CPU oracles (each pinned separately), tiny widths, 4 DDIM steps, mild guidance (a large CFG
scale multiplies the bf16 error of y - u by ~2s). <= 10 % inf / 8 % L2 on the decoded frames."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_image_to_video_chain_vs_oracle_chain(cuda):
    from oracle import restate_openclip as OC, restate_unet as U, restate_vae as V
    from oracle.weights import seeded_state_dict
    from vitron_b200 import param_shapes
    from vitron_b200.autoencoder import AutoencoderKL
    from vitron_b200.clip_embedder import FrozenOpenCLIPTtxtVisualEmbedder
    from vitron_b200.i2vgen_pipeline import I2VGenXLPipeline
    from vitron_b200.unet_i2vgen import UNetSD_I2VGen
    ufx = torch.load(os.path.join(GOLD, "unet_tiny.pt"), weights_only=False)
    vfx = torch.load(os.path.join(GOLD, "vae_tiny.pt"), weights_only=False)
    usd = seeded_state_dict(ufx["shapes"], ufx["seed"], ufx["gain"])
    vsd = seeded_state_dict(vfx["shapes"], vfx["seed"], 0.8)
    ccfg = dict(embed_dim=1024, text=dict(width=1024, layers=1, heads=16, context_length=77, vocab_size=512),
                vision=dict(width=128, layers=1, heads=2, patch_size=14, image_size=56, mlp=256))
    csd = seeded_state_dict(OC.openclip_shapes(ccfg), 23)
    F_, scale, steps, guide = 4, 0.18215, 4, 1.5
    g = torch.Generator().manual_seed(12)
    img_vit, img_vae = torch.randn((1, 3, 56, 56), generator=g), torch.randn((1, 3, 32, 64), generator=g)
    tokens = torch.randint(1, 511, (1, 77), generator=g)
    tokens[0, 20], tokens[0, 21:] = 511, 0
    neg = torch.randint(1, 511, (1, 77), generator=g)
    neg[0, 9], neg[0, 10:] = 511, 0
    post_noise = torch.randn((1, 4, 8, 16), generator=g)
    noise = torch.randn((1, 4, F_, 8, 16), generator=g)

    def oracle_chain(img_vit, img_vae, tokens, neg, post_noise, noise):
        _, y_words = OC.encode_text(csd, tokens, ccfg, layer_idx=1)
        _, y_neg = OC.encode_text(csd, neg, ccfg, layer_idx=1)
        y_vis = OC.encode_image(csd, img_vit, ccfg).unsqueeze(1)
        mean, _, std = V.encode_moments(vsd, img_vae, vfx["ddconfig"])
        local = (scale * (mean + std * post_noise)).unsqueeze(2).repeat_interleave(F_, dim=2)
        fps = torch.tensor([16])
        kw = [dict(y=y_words, image=y_vis, local_image=local, fps=fps),
              dict(y=y_neg, image=torch.zeros_like(y_vis), local_image=local, fps=fps)]
        omodel = lambda xt, t, **k: U.unet_forward(usd, ufx["cfg"], xt, t, **k)
        lat = U.ddim_sample_loop(noise, omodel, kw, guide, steps) / scale
        ref = V.decode(vsd, lat.permute(0, 2, 1, 3, 4).reshape(F_, 4, 8, 16), vfx["ddconfig"])
        return ref.reshape(1, F_, 3, 32, 64).permute(0, 2, 1, 3, 4)

    unet = UNetSD_I2VGen(**ufx["cfg"], device=cuda)
    unet.load_state_dict(usd)
    vae = AutoencoderKL(vfx["ddconfig"], 4, device=cuda).load_state_dict(vsd)
    clip = FrozenOpenCLIPTtxtVisualEmbedder(None, device=cuda, layer="penultimate", arch_cfg=ccfg).load_state_dict(csd)
    pipe = I2VGenXLPipeline(unet, vae, clip, scale_factor=scale, max_frames=F_, guide_scale=guide, ddim_timesteps=steps, decoder_bs=2)

    def check(got, ref, what):
        assert tuple(got.shape) == (1, 3, F_, 32, 64) and bool(torch.isfinite(got).all())
        got, r = got.float().cpu(), ref.float()
        e_inf = ((got - r).abs().max() / (r.abs().max() + 1e-6)).item()
        e_l2 = ((got - r).norm() / (r.norm() + 1e-6)).item()
        assert e_inf < 0.10 and e_l2 < 0.08, (what, e_inf, e_l2)

    a1 = (img_vit, img_vae, tokens, neg, post_noise, noise)
    check(pipe(img_vit, img_vae, tokens, neg, noise=noise, posterior_noise=post_noise), oracle_chain(*a1), "first video (graph capture)")
    # second video: different image / prompts / noise through the SAME captured graph (conditioning rebound in place)
    g2 = torch.Generator().manual_seed(77)
    img_vit2, img_vae2 = torch.randn((1, 3, 56, 56), generator=g2), torch.randn((1, 3, 32, 64), generator=g2)
    tokens2 = torch.randint(1, 511, (1, 77), generator=g2)
    tokens2[0, 33], tokens2[0, 34:] = 511, 0
    post2, noise2 = torch.randn((1, 4, 8, 16), generator=g2), torch.randn((1, 4, F_, 8, 16), generator=g2)
    den = pipe._den
    v2 = pipe(img_vit2, img_vae2, tokens2, neg, noise=noise2, posterior_noise=post2)
    assert pipe._den is den, "the captured graph must be reused"
    check(v2, oracle_chain(img_vit2, img_vae2, tokens2, neg, post2, noise2), "second video (rebound graph)")
