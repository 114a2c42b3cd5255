"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of i2vgen-xl's UNetSD_I2VGen forward and of the
DDIM / classifier-free-guidance sampler.  Plain torch; never imported by the product package.

Follows modules/i2vgen-xl/tools/modules/unet/unet_i2vgen.py:243-418 and util.py (ResBlock :610-730,
TemporalConvBlock_v2 :1347-1392, SpatialTransformer :311-373, BasicTransformerBlock :510-540,
GEGLU :543-549, FeedForward :560-577, TemporalTransformer :992-1089, TransformerV2 :1129-1148,
Upsample :579-607, Downsample :732-756, sinusoidal_embedding :177-189) and
tools/modules/diffusions/diffusion_ddim.py:143-250, schedules.py:50-57,121-143.
xformers.memory_efficient_attention is restated as softmax(QK^T/sqrt(d))V (its definition).

Parity status: PINNED by tests/test_oracle_cpu.py against tests/golden/unet_tiny.pt (generated from
the unmodified reference by oracle/gen_golden.py) and against the live reference when present.
"""
import math

import torch
import torch.nn.functional as F


def sinusoidal_embedding(timesteps, dim):
    half = dim // 2
    timesteps = timesteps.float()
    sinusoid = torch.outer(timesteps, torch.pow(10000, -torch.arange(half).to(timesteps).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def plan(cfg):
    """Block list with state-dict prefixes (unet_i2vgen.py:133-233)."""
    dim, dim_mult, nres = cfg["dim"], list(cfg["dim_mult"]), cfg["num_res_blocks"]
    hd, scales = cfg["head_dim"], list(cfg["attn_scales"])
    heads0 = cfg.get("num_heads") or dim // 32
    enc = [dim * u for u in [1] + dim_mult]
    dec = [dim * u for u in [dim_mult[-1]] + dim_mult[::-1]]
    sc, scale = [dim], 1.0
    inp = [[("conv_in", "input_blocks.0.0", 0, 0), ("tt", "input_blocks.0.1", dim, heads0)]]
    idx = 1
    for i, (ci, co) in enumerate(zip(enc[:-1], enc[1:])):
        for j in range(nres):
            b = [("res", f"input_blocks.{idx}.0", ci, co)]
            if scale in scales:
                b += [("st", f"input_blocks.{idx}.1", co, co // hd), ("tt", f"input_blocks.{idx}.2", co, co // hd)]
            ci = co
            inp.append(b)
            sc.append(co)
            idx += 1
            if i != len(dim_mult) - 1 and j == nres - 1:
                inp.append([("down", f"input_blocks.{idx}", co, co)])
                sc.append(co)
                scale /= 2
                idx += 1
    c = enc[-1]
    mid = [("res", "middle_block.0", c, c), ("st", "middle_block.1", c, c // hd), ("tt", "middle_block.2", c, c // hd),
           ("res", "middle_block.3", c, c)]
    outp, idx = [], 0
    for i, (ci, co) in enumerate(zip(dec[:-1], dec[1:])):
        for j in range(nres + 1):
            b = [("res", f"output_blocks.{idx}.0", ci + sc.pop(), co)]
            k = 1
            if scale in scales:
                b += [("st", f"output_blocks.{idx}.1", co, co // hd), ("tt", f"output_blocks.{idx}.2", co, co // hd)]
                k = 3
            ci = co
            if i != len(dim_mult) - 1 and j == nres:
                b.append(("up", f"output_blocks.{idx}.{k}", co, co))
                scale *= 2
            outp.append(b)
            idx += 1
    return inp, mid, outp


def _attn(x, ctx, sd, p, heads):
    """MemoryEfficientCrossAttention (util.py:212-267)."""
    ctx = x if ctx is None else ctx
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(ctx, sd[p + ".to_k.weight"])
    v = F.linear(ctx, sd[p + ".to_v.weight"])
    b, n, inner = q.shape
    hd = inner // heads
    q, k, v = (t.view(b, t.shape[1], heads, hd).transpose(1, 2) for t in (q, k, v))
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), -1) @ v
    a = a.transpose(1, 2).reshape(b, n, inner)
    return F.linear(a, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _tblock(x, ctx, sd, p, heads):
    ln = lambda t, n: F.layer_norm(t, (t.shape[-1],), sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"])
    x = _attn(ln(x, "norm1"), None, sd, p + ".attn1", heads) + x
    x = _attn(ln(x, "norm2"), ctx, sd, p + ".attn2", heads) + x
    h = F.linear(ln(x, "norm3"), sd[p + ".ff.net.0.proj.weight"], sd[p + ".ff.net.0.proj.bias"])
    a, gate = h.chunk(2, dim=-1)
    return F.linear(a * F.gelu(gate), sd[p + ".ff.net.2.weight"], sd[p + ".ff.net.2.bias"]) + x


def _gn(x, sd, p, eps=1e-5):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _res(x, emb, sd, p, batch):
    h = F.conv2d(F.silu(_gn(x, sd, p + ".in_layers.0")), sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.conv2d(F.silu(_gn(h, sd, p + ".out_layers.0")), sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    skip = x if (p + ".skip_connection.weight") not in sd else F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    h = skip + h
    bf, c, hh, ww = h.shape
    t = h.view(batch, bf // batch, c, hh, ww).permute(0, 2, 1, 3, 4)
    ident = t
    for ci, li in ((1, 2), (2, 3), (3, 3), (4, 3)):
        q = f"{p}.temopral_conv.conv{ci}"
        t = F.conv3d(F.silu(_gn(t, sd, q + ".0")), sd[f"{q}.{li}.weight"], sd[f"{q}.{li}.bias"], padding=(1, 0, 0))
    t = ident + t
    return t.permute(0, 2, 1, 3, 4).reshape(bf, c, hh, ww)


def _st(x, context, sd, p, heads):
    b, c, h, w = x.shape
    z = _gn(x, sd, p + ".norm", 1e-6).flatten(2).transpose(1, 2)
    z = F.linear(z, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    z = _tblock(z, context, sd, p + ".transformer_blocks.0", heads)
    z = F.linear(z, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return z.transpose(1, 2).reshape(b, c, h, w) + x


def _tt(x, sd, p, heads, batch):
    bf, c, h, w = x.shape
    f = bf // batch
    x5 = x.view(batch, f, c, h, w).permute(0, 2, 1, 3, 4)  # b c f h w
    z = _gn(x5, sd, p + ".norm", 1e-6)
    z = z.permute(0, 3, 4, 1, 2).reshape(batch * h * w, c, f)
    z = F.conv1d(z, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"]).transpose(1, 2)  # bhw f inner
    z = _tblock(z, None, sd, p + ".transformer_blocks.0", heads)
    z = F.conv1d(z.transpose(1, 2), sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])  # bhw c f
    z = z.view(batch, h, w, c, f).permute(0, 3, 4, 1, 2)
    out = z + x5
    return out.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


def unet_forward(sd, cfg, x, t, y=None, image=None, local_image=None, fps=None, zero_y=None):
    """UNetSD_I2VGen.forward (inference path). x [b,c,f,h,w] -> [b,out,f,h,w] fp32."""
    sd = {k: v.float() for k, v in sd.items()}
    batch, c, f, h, w = x.shape
    dim = cfg["dim"]
    if local_image.ndim == 5 and local_image.size(2) > 1:
        local_image = local_image[:, :, :1]
    elif local_image.ndim != 5:
        local_image = local_image.unsqueeze(2)
    local_image = local_image.float()
    # [Concat] (:278-295)
    if f > 1:
        mask_pos = torch.cat([torch.ones_like(local_image[:, :, :1]) * ((tp + 1) / (f - 1)) for tp in range(f - 1)], dim=2)
        ximg = torch.cat([local_image[:, :, :1], mask_pos], dim=2)
    else:
        ximg = local_image
    ximg = ximg.permute(0, 2, 1, 3, 4).reshape(batch * f, -1, h, w)
    for i, k in enumerate((0, 2, 4)):
        if i:
            ximg = F.silu(ximg)
        ximg = F.conv2d(ximg, sd[f"local_image_concat.{k}.weight"], sd[f"local_image_concat.{k}.bias"], padding=1)
    cd = ximg.shape[1]
    s = ximg.view(batch, f, cd, h, w).permute(0, 3, 4, 1, 2).reshape(batch * h * w, f, cd)
    li = 0
    while f"local_temporal_encoder.layers.{li}.0.norm.weight" in sd:
        p = f"local_temporal_encoder.layers.{li}."
        yv = F.layer_norm(s, (cd,), sd[p + "0.norm.weight"], sd[p + "0.norm.bias"])
        qkv = F.linear(yv, sd[p + "0.fn.to_qkv.weight"]).chunk(3, dim=-1)
        q, k, v = (u.view(u.shape[0], f, 2, cd).transpose(1, 2) for u in qkv)
        a = torch.softmax(q @ k.transpose(-1, -2) * cd ** -0.5, -1) @ v
        a = a.transpose(1, 2).reshape(s.shape[0], f, 2 * cd)
        s = F.linear(a, sd[p + "0.fn.to_out.0.weight"], sd[p + "0.fn.to_out.0.bias"]) + s
        s = F.linear(F.gelu(F.linear(s, sd[p + "1.net.0.0.weight"], sd[p + "1.net.0.0.bias"])), sd[p + "1.net.2.weight"], sd[p + "1.net.2.bias"]) + s
        li += 1
    ximg5 = s.view(batch, h, w, f, cd).permute(0, 4, 3, 1, 2)
    concat = ximg5 + ximg5
    # [Embeddings]
    mlp = lambda n, v: F.linear(F.silu(F.linear(v, sd[n + ".0.weight"], sd[n + ".0.bias"])), sd[n + ".2.weight"], sd[n + ".2.bias"])
    emb = mlp("time_embed", sinusoidal_embedding(t, dim)) + mlp("fps_embedding", sinusoidal_embedding(fps, dim))
    emb = emb.repeat_interleave(repeats=f, dim=0)
    # [Context]
    ctx = [y.float()] if y is not None else [zero_y.repeat(batch, 1, 1)[:, :1].float()]
    lc = local_image[:, :, 0]
    lc = F.silu(F.conv2d(lc, sd["local_image_embedding.0.weight"], sd["local_image_embedding.0.bias"], padding=1))
    lc = F.adaptive_avg_pool2d(lc, (32, 32))
    lc = F.silu(F.conv2d(lc, sd["local_image_embedding.3.weight"], sd["local_image_embedding.3.bias"], stride=2, padding=1))
    lc = F.conv2d(lc, sd["local_image_embedding.5.weight"], sd["local_image_embedding.5.bias"], stride=2, padding=1)
    ctx.append(lc.flatten(2).transpose(1, 2))
    if image is not None:
        ctx.append(mlp("context_embedding", image.float()).view(-1, cfg.get("num_tokens", 4), cfg["context_dim"]))
    context = torch.cat(ctx, dim=1).repeat_interleave(repeats=f, dim=0)

    xx = torch.cat([x.float(), concat], dim=1).permute(0, 2, 1, 3, 4).reshape(batch * f, -1, h, w)
    inp, mid, outp = plan(cfg)

    def run(blk, xx):
        for kind, p, ci, co in blk:
            if kind == "conv_in":
                xx = F.conv2d(xx, sd[p + ".weight"], sd[p + ".bias"], padding=1)
            elif kind == "res":
                xx = _res(xx, emb, sd, p, batch)
            elif kind == "st":
                xx = _st(xx, context, sd, p, co)
            elif kind == "tt":
                xx = _tt(xx, sd, p, co, batch)
            elif kind == "down":
                xx = F.conv2d(xx, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
            elif kind == "up":
                xx = F.conv2d(F.interpolate(xx, scale_factor=2, mode="nearest"), sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
        return xx

    xs = []
    for blk in inp:
        xx = run(blk, xx)
        xs.append(xx)
    xx = run(mid, xx)
    for blk in outp:
        xx = run(blk, torch.cat([xx, xs.pop()], dim=1))
    xx = F.conv2d(F.silu(_gn(xx, sd, "out.0")), sd["out.2.weight"], sd["out.2.bias"], padding=1)
    return xx.view(batch, f, -1, h, w).permute(0, 2, 1, 3, 4)


# ------------------------------------------------------------------ DDIM
def cosine_betas(num_timesteps=1000, cosine_s=0.008, zero_terminal_snr=True):
    fn = lambda u: math.cos((u + cosine_s) / (1 + cosine_s) * math.pi / 2) ** 2
    betas = torch.tensor([min(1.0 - fn((s + 1) / num_timesteps) / fn(s / num_timesteps), 0.999) for s in range(num_timesteps)],
                         dtype=torch.float64)
    if zero_terminal_snr and betas.max() != 1.0:
        a = (1 - betas).cumprod(0).sqrt()
        a0, aT = a[0].clone(), a[-1].clone()
        a = (a - aT) * a0 / (a0 - aT)
        ab = a ** 2
        betas = 1 - torch.cat([ab[0:1], ab[1:] / ab[:-1]])
    return betas


def ddim_sample_loop(noise, model, model_kwargs, guide_scale, ddim_timesteps, betas=None, mean_type="v"):
    """eta = 0 DDIM with classifier-free guidance; model(xt, t, **kw) -> prediction."""
    betas = cosine_betas() if betas is None else betas
    T = len(betas)
    ac = torch.cumprod(1 - betas, 0)
    _i = lambda tensor, t, x: tensor[t].view(x.size(0), *((1,) * (x.ndim - 1))).to(x)
    stride = T // ddim_timesteps
    xt = noise.float()
    b = noise.size(0)
    steps = (1 + torch.arange(0, T, stride)).clamp(0, T - 1).flip(0)
    for step in steps:
        t = torch.full((b,), int(step), dtype=torch.long)
        y_out = model(xt, t, **model_kwargs[0])
        u_out = model(xt, t, **model_kwargs[1])
        out = u_out + guide_scale * (y_out - u_out)
        if mean_type == "v":
            x0 = _i(ac.sqrt(), t, xt) * xt - _i((1 - ac).sqrt(), t, xt) * out
        else:
            x0 = _i((1 / ac).sqrt(), t, xt) * xt - _i((1 / ac - 1).sqrt(), t, xt) * out
        eps = (_i((1 / ac).sqrt(), t, xt) * xt - x0) / _i((1 / ac - 1).sqrt(), t, xt)
        a_prev = _i(ac, (t - stride).clamp(0), xt)
        xt = torch.sqrt(a_prev) * x0 + torch.sqrt(1 - a_prev) * eps
    return xt
