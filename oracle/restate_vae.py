"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of i2vgen-xl's first-stage AutoencoderKL.
Plain torch functional code over a state dict; never imported by the product package.

Follows modules/i2vgen-xl/tools/modules/autoencoder.py: Normalize = GroupNorm(32, eps=1e-6) :15-16,
nonlinearity = x*sigmoid(x) :11-13, ResnetBlock.forward :315-336 (temb is None), AttnBlock.forward :418-442
(single head of `c` channels, scale c^-0.5), Upsample :455-459 (nearest x2 + conv3x3), Downsample :474-481
(zero pad right/bottom by 1, conv3x3 stride 2 pad 0), Encoder.forward :549-578, Decoder.forward :653-686,
AutoencoderKL.encode :79-83 / decode :100-103, DiagonalGaussianDistribution :212-253 (mean, logvar clamped to
[-30, 20]); call sites tools/inferences/inference_i2vgen_entrance.py:172-173 (`encode_firsr_stage`), :205-208
(`decode` of 1/scale_factor * latents in chunks of decoder_bs frames).

Parity status: PINNED — tests/golden/vae_tiny.pt comes from the unmodified AutoencoderKL class (oracle/gen_golden.py::
gen_vae) and tests/test_oracle_cpu.py::test_vae_restatement_matches_* compare this file against it and the live class.
"""
import torch
import torch.nn.functional as F

SD_VAE = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 4, 4),
              num_res_blocks=2, attn_resolutions=(), dropout=0.0)
"""tools/modules/config.py:110-127 (the Stable-Diffusion 2.1 VAE)."""


def _gn_silu(x, sd, p):
    x = F.group_norm(x, 32, sd[p + "weight"].float(), sd[p + "bias"].float(), 1e-6)
    return x * torch.sigmoid(x)


def _conv(x, sd, p, stride=1, padding=1):
    return F.conv2d(x, sd[p + "weight"].float(), sd[p + "bias"].float(), stride=stride, padding=padding)


def resnet_block(x, sd, p):
    h = _conv(_gn_silu(x, sd, p + "norm1."), sd, p + "conv1.")
    h = _conv(_gn_silu(h, sd, p + "norm2."), sd, p + "conv2.")
    if p + "nin_shortcut.weight" in sd:
        x = _conv(x, sd, p + "nin_shortcut.", padding=0)
    elif p + "conv_shortcut.weight" in sd:
        x = _conv(x, sd, p + "conv_shortcut.")
    return x + h


def attn_block(x, sd, p):
    h = F.group_norm(x, 32, sd[p + "norm.weight"].float(), sd[p + "norm.bias"].float(), 1e-6)
    q, k, v = (_conv(h, sd, p + n, padding=0) for n in ("q.", "k.", "v."))
    b, c, hh, ww = q.shape
    w_ = torch.bmm(q.reshape(b, c, hh * ww).permute(0, 2, 1), k.reshape(b, c, hh * ww)) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    h = torch.bmm(v.reshape(b, c, hh * ww), w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(h, sd, p + "proj_out.", padding=0)


def encoder_forward(sd, x, cfg=SD_VAE, p="encoder."):
    nres = len(cfg["ch_mult"])
    h = _conv(x.float(), sd, p + "conv_in.")
    for i in range(nres):
        for j in range(cfg["num_res_blocks"]):
            h = resnet_block(h, sd, p + f"down.{i}.block.{j}.")
            if p + f"down.{i}.attn.{j}.norm.weight" in sd:
                h = attn_block(h, sd, p + f"down.{i}.attn.{j}.")
        if i != nres - 1:
            h = _conv(F.pad(h, (0, 1, 0, 1)), sd, p + f"down.{i}.downsample.conv.", stride=2, padding=0)
    h = resnet_block(h, sd, p + "mid.block_1.")
    h = attn_block(h, sd, p + "mid.attn_1.")
    h = resnet_block(h, sd, p + "mid.block_2.")
    return _conv(_gn_silu(h, sd, p + "norm_out."), sd, p + "conv_out.")


def decoder_forward(sd, z, cfg=SD_VAE, p="decoder."):
    nres = len(cfg["ch_mult"])
    h = _conv(z.float(), sd, p + "conv_in.")
    h = resnet_block(h, sd, p + "mid.block_1.")
    h = attn_block(h, sd, p + "mid.attn_1.")
    h = resnet_block(h, sd, p + "mid.block_2.")
    for i in reversed(range(nres)):
        for j in range(cfg["num_res_blocks"] + 1):
            h = resnet_block(h, sd, p + f"up.{i}.block.{j}.")
            if p + f"up.{i}.attn.{j}.norm.weight" in sd:
                h = attn_block(h, sd, p + f"up.{i}.attn.{j}.")
        if i != 0:
            h = _conv(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd, p + f"up.{i}.upsample.conv.")
    return _conv(_gn_silu(h, sd, p + "norm_out."), sd, p + "conv_out.")


def encode_moments(sd, x, cfg=SD_VAE):
    """AutoencoderKL.encode: (mean, logvar clamped, std) of the posterior."""
    m = _conv(encoder_forward(sd, x, cfg), sd, "quant_conv.", padding=0)
    mean, logvar = torch.chunk(m, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean, logvar, torch.exp(0.5 * logvar)


def decode(sd, z, cfg=SD_VAE):
    return decoder_forward(sd, _conv(z.float(), sd, "post_quant_conv.", padding=0), cfg)
