"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of GLIGEN's grounded SD UNet; plain torch
functional code over a state dict, never imported by the product package.

Follows modules/GLIGEN/demo/gligen/ldm/modules/diffusionmodules/openaimodel.py: UNetModel.__init__ block order :283-372,
forward :447-502, forward_position_net :388-405, ResBlock._forward :209-230 (use_scale_shift_norm=False),
Downsample :84-110 (conv k3 s2 p1), Upsample :51-81 (nearest x2 + conv), TimestepEmbedSequential :34-48;
ldm/modules/attention.py SpatialTransformer :352-386 (GroupNorm eps 1e-6, 1x1 proj_in / proj_out, residual);
diffusionmodules/positionnet.py:30-50; util.py timestep_embedding :160-180, FourierEmbedder :12-26, GroupNorm32 :223-225.
Transformer blocks: oracle/restate_gligen.py::basic_transformer_block (pinned separately).

Parity status: PINNED — tests/golden/gligen_unet_tiny.pt is produced by the unmodified UNetModel class
(oracle/gen_golden.py::gen_gligen_unet); tests/test_oracle_cpu.py::test_gligen_unet_restatement_matches_* compare.
"""
import math

import torch
import torch.nn.functional as F

from .restate_gligen import basic_transformer_block


def timestep_embedding(timesteps, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def fourier_embed(x, num_freqs=8, temperature=100):
    bands = temperature ** (torch.arange(num_freqs) / num_freqs)
    out = []
    for f in bands:
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


def _conv(x, sd, p, stride=1, padding=1):
    return F.conv2d(x, sd[p + "weight"].float(), sd[p + "bias"].float(), stride=stride, padding=padding)


def _gn(x, sd, p, eps=1e-5):
    return F.group_norm(x, 32, sd[p + "weight"].float(), sd[p + "bias"].float(), eps)


def position_net(sd, boxes, masks, text, p="position_net."):
    masks = masks.unsqueeze(-1).float()
    xyxy = fourier_embed(boxes.float())
    text = text.float() * masks + (1 - masks) * sd[p + "null_positive_feature"].float().view(1, 1, -1)
    xyxy = xyxy * masks + (1 - masks) * sd[p + "null_position_feature"].float().view(1, 1, -1)
    h = torch.cat([text, xyxy], dim=-1)
    h = F.silu(F.linear(h, sd[p + "linears.0.weight"].float(), sd[p + "linears.0.bias"].float()))
    h = F.silu(F.linear(h, sd[p + "linears.2.weight"].float(), sd[p + "linears.2.bias"].float()))
    return F.linear(h, sd[p + "linears.4.weight"].float(), sd[p + "linears.4.bias"].float())


def res_block(x, emb, sd, p):
    h = _conv(F.silu(_gn(x, sd, p + "in_layers.0.")), sd, p + "in_layers.2.")
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"].float(), sd[p + "emb_layers.1.bias"].float())
    h = h + e[:, :, None, None]
    h = _conv(F.silu(_gn(h, sd, p + "out_layers.0.")), sd, p + "out_layers.3.")
    if p + "skip_connection.weight" in sd:
        x = _conv(x, sd, p + "skip_connection.", padding=0)
    return x + h


def spatial_transformer(x, context, objs, sd, p, heads, depth, scale=1.0):
    b, c, h, w = x.shape
    t = _conv(_gn(x, sd, p + "norm.", 1e-6), sd, p + "proj_in.", padding=0)
    t = t.flatten(2).transpose(1, 2)
    for d in range(depth):
        t = basic_transformer_block(sd, p + f"transformer_blocks.{d}.", t, context, objs, heads, scale)
    t = t.transpose(1, 2).reshape(b, -1, h, w)
    return _conv(t, sd, p + "proj_out.", padding=0) + x


def unet_forward(sd, cfg, inp, alpha_scale=1.0):
    """cfg: UNetModel constructor kwargs; inp: the reference's input dict -> [B, out_channels, H, W].
    alpha_scale = the `scale` evaluator.py::set_alpha_scale puts on every gated fuser (the sampler's gate schedule)."""
    mc, mult, nrb = cfg["model_channels"], tuple(cfg["channel_mult"]), cfg["num_res_blocks"]
    att, heads, depth = list(cfg["attention_resolutions"]), cfg.get("num_heads", 8), cfg.get("transformer_depth", 1)
    x = inp["x"].float()
    if "boxes" in inp:
        objs = position_net(sd, inp["boxes"], inp["masks"], inp["text_embeddings"])
    else:
        b, n = x.shape[0], cfg.get("max_box", 30)
        objs = position_net(sd, torch.zeros(b, n, 4), torch.zeros(b, n), torch.zeros(b, n, cfg.get("positive_len", 768)))
    emb = timestep_embedding(inp["timesteps"], mc)
    emb = F.linear(F.silu(F.linear(emb, sd["time_embed.0.weight"].float(), sd["time_embed.0.bias"].float())),
                   sd["time_embed.2.weight"].float(), sd["time_embed.2.bias"].float())
    if cfg.get("is_inpaint", False):
        x = torch.cat([x, inp["inpainting_extra_input"].float()], dim=1)
    context = inp["context"].float()
    st = lambda t, p: spatial_transformer(t, context, objs, sd, p, heads, depth, alpha_scale)
    hs = []
    h = _conv(x, sd, "input_blocks.0.0.")
    hs.append(h)
    ds, idx = 1, 1
    for level in range(len(mult)):
        for _ in range(nrb):
            h = res_block(h, emb, sd, f"input_blocks.{idx}.0.")
            if ds in att:
                h = st(h, f"input_blocks.{idx}.1.")
            hs.append(h)
            idx += 1
        if level != len(mult) - 1:
            h = _conv(h, sd, f"input_blocks.{idx}.0.op.", stride=2)
            hs.append(h)
            ds *= 2
            idx += 1
    h = res_block(h, emb, sd, "middle_block.0.")
    h = st(h, "middle_block.1.")
    h = res_block(h, emb, sd, "middle_block.2.")
    idx = 0
    for level in reversed(range(len(mult))):
        for i in range(nrb + 1):
            h = res_block(torch.cat([h, hs.pop()], dim=1), emb, sd, f"output_blocks.{idx}.0.")
            j = 1
            if ds in att:
                h = st(h, f"output_blocks.{idx}.{j}.")
                j += 1
            if level and i == nrb:
                h = _conv(F.interpolate(h, scale_factor=2, mode="nearest"), sd, f"output_blocks.{idx}.{j}.conv.")
                ds //= 2
            idx += 1
    return _conv(F.silu(_gn(h, sd, "out.0.")), sd, "out.2.")
