"""TEST INFRASTRUCTURE ONLY — torch (CPU) statements of the vitron_b200.ops entry points used by the host-side
modules, so that their HOST LOGIC (weight folding / packing, view slicing, call order, shapes) can be checked
without a GPU (`-m "not gpu"`). Never imported by the product: tests install it with monkeypatch over
`vitron_b200.ops`; the kernels themselves are checked on the GPU against the oracle (tests/*_gpu.py).
Each function mirrors the contract in vitron_b200/ops.py: bf16 storage, fp32 arithmetic."""
import math

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16
ACT_NONE, ACT_GELU, ACT_QUICK_GELU, ACT_RELU, ACT_SILU = 0, 1, 2, 3, 4


def _act(x, act):
    return {ACT_NONE: lambda t: t, ACT_GELU: F.gelu, ACT_QUICK_GELU: lambda t: t * torch.sigmoid(1.702 * t),
            ACT_RELU: F.relu, ACT_SILU: F.silu}[int(act)](x)


def _unglu(v):
    """Columns of a GEMM against an ops.pack_glu_weight matrix: 16-column blocks alternate a | b."""
    n = v.shape[-1]
    blk = v.reshape(*v.shape[:-1], n // 32, 2, 16)
    return blk[..., 0, :].reshape(*v.shape[:-1], n // 2), blk[..., 1, :].reshape(*v.shape[:-1], n // 2)


def gemm(a, w, bias=None, act=ACT_NONE, glu=0, residual=None, alpha=1.0, rowbias=None, rowbias_rows=0, out=None,
         out_fp32=False, rowscale=None, rms_eps=0.0):
    assert a.dtype == BF16 and w.dtype == BF16 and a.shape[-1] == w.shape[1]
    a2 = a.float().reshape(-1, a.shape[-1])
    v = a2 @ w.float().t()
    if rowscale is not None:
        v = v * rowscale.float().reshape(-1, 1)
    elif rms_eps:                                   # folded RMSNorm: rows scaled by rsqrt(mean(a^2) + eps)
        v = v * torch.rsqrt(a2.pow(2).mean(-1, keepdim=True) + rms_eps)
    if bias is not None:
        v = v + bias.float()
    if rowbias is not None:
        v = v + rowbias.float().repeat_interleave(int(rowbias_rows), 0)[:v.shape[0]]
    if glu:
        x, g = _unglu(v)
        v = F.silu(x) * g if glu == 1 else x * F.gelu(g)
    else:
        v = _act(v, act)
    v = v * alpha
    if residual is not None:
        v = residual.float().reshape(v.shape) + v
    v = v.reshape(*a.shape[:-1], v.shape[-1]).to(torch.float32 if out_fp32 else BF16)
    if out is not None:
        out.copy_(v.reshape(out.shape))
        return out
    return v


def pack_glu_weight(w_a, w_b):
    f = w_a.shape[0]
    rest = w_a.shape[1:]
    return torch.stack([w_a.reshape(f // 16, 16, *rest), w_b.reshape(f // 16, 16, *rest)], dim=1).reshape(2 * f, *rest).contiguous()


def attention(q, k, v, scale=None, causal=False, kv_len=None, mask=None, out=None):
    B, Sq, H, D = q.shape
    scale = D ** -0.5 if scale is None else scale
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
    if causal:
        Skv = k.shape[1]
        s = s.masked_fill(torch.ones(Sq, Skv, dtype=torch.bool).triu(Skv - Sq + 1), float("-inf"))
    if mask is not None:
        s = s.masked_fill(mask.bool(), float("-inf"))
    if kv_len is not None:
        dead = torch.arange(k.shape[1])[None, :] >= kv_len.long().reshape(-1, 1)
        s = s.masked_fill(dead[:, None, None, :], float("-inf"))
    o = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1).nan_to_num(), v.float()).to(BF16).contiguous()
    if out is not None:
        out.copy_(o)
        return out
    return o


def layernorm(x, weight, bias, eps, out=None):
    v = F.layer_norm(x.float(), (x.shape[-1],), weight.float(), None if bias is None else bias.float(), eps).to(BF16)
    if out is not None:
        out.copy_(v)
        return out
    return v


def layernorm_add(x, weight, bias, residual, eps, out=None):
    v = F.layer_norm(x.float(), (x.shape[-1],), weight.float(), None if bias is None else bias.float(), eps)
    if residual is not None:
        v = v + residual.float()
    v = v.to(BF16)
    if out is not None:
        out.copy_(v)
        return out
    return v


def pack_dwconv_weight(w):
    c, one, kh, kw = w.shape
    return w.reshape(c, kh * kw).t().to(BF16).contiguous()


def dwconv_nhwc(x, wt, k, act=ACT_NONE):
    nb, h, w, c = x.shape
    assert x.stride(3) == 1 and x.stride(1) == w * x.stride(2) and x.stride(0) == h * w * x.stride(2)
    assert wt.shape == (k * k, c)
    wc = wt.float().t().reshape(c, 1, k, k)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wc, None, padding=k // 2, groups=c)
    return _act(y, act).permute(0, 2, 3, 1).to(BF16).contiguous()


def colmean(x, nb, act=ACT_NONE):
    c = x.shape[-1]
    return _act(x.float().reshape(nb, -1, c).mean(1), act)


def focal_modulate(levels, gates, glob, nb, scale):
    c = levels[0].shape[-1]
    t = levels[0].numel() // (nb * c)
    g = gates.float().reshape(nb * t, -1)
    v = glob.repeat_interleave(t, 0) * g[:, len(levels):len(levels) + 1]
    for l, lv in enumerate(levels):
        v = v + lv.float().reshape(nb * t, c) * g[:, l:l + 1]
    return (v * scale).to(BF16)


def mul_rows(a, b):
    return (a.float() * b.float()).to(BF16)


def im2col_nchw(pixels, k, stride, pad, ho, wo, kpad):
    nb, c, h, w = pixels.shape
    hp, wp = (ho - 1) * stride + k - 2 * pad, (wo - 1) * stride + k - 2 * pad  # extent the output grid reads
    xp = F.pad(pixels.float(), (0, max(0, wp - w), 0, max(0, hp - h)))
    cols = F.unfold(xp, k, padding=pad, stride=stride).transpose(1, 2)[:, :ho * wo].reshape(nb * ho * wo, c * k * k)
    out = torch.zeros((nb * ho * wo, kpad), dtype=BF16)
    out[:, :c * k * k] = cols.to(BF16)
    return out


def pack_conv_weight(w):
    if w.dim() == 5:
        w = w[:, :, :, 0, 0].unsqueeze(-1)
    cout, cin, kh, kw = w.shape
    cpad = (cin + 63) // 64 * 64
    out = torch.zeros((cout, kh * kw, cpad), dtype=BF16)
    out[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin).to(BF16)
    return out


def conv_nhwc(x, wt, kh, kw, stride=1, pad_h=None, pad_w=None, bias=None, act=ACT_NONE, glu=0, residual=None, alpha=1.0,
              rowbias=None, rowbias_rows=0, out=None):
    assert glu == 0 and out is None
    nb, h, w, cin = x.shape
    cout = wt.shape[0]
    pad_h = kh // 2 if pad_h is None else pad_h
    pad_w = kw // 2 if pad_w is None else pad_w
    wc = wt.float()[:, :, :cin].reshape(cout, kh, kw, cin).permute(0, 3, 1, 2)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wc, None if bias is None else bias.float(), stride=stride,
                 padding=(pad_h, pad_w))
    y = y.permute(0, 2, 3, 1)
    if rowbias is not None:  # per (rowbias_rows consecutive output pixels) group bias over the channels
        rb = rowbias.float().repeat_interleave(int(rowbias_rows), 0)[:y.numel() // y.shape[-1]]
        y = y + rb.reshape(y.shape)
    y = _act(y, act) * alpha
    if residual is not None:
        y = residual.float() + y
    return y.to(BF16).contiguous()


def groupnorm_nhwc(x, weight, bias, groups, eps, act=ACT_NONE, n=None, out=None):
    c = x.shape[-1]
    n = x.shape[0] if n is None else n
    y = F.group_norm(x.float().reshape(n, -1, c).permute(0, 2, 1), groups, weight.float(), bias.float(), eps)
    return _act(y, act).permute(0, 2, 1).reshape(x.shape).to(BF16).contiguous()


def conv_nhwc_direct(x, w_khwc, bias, kh, kw, stride=1, pad_h=None, pad_w=None):
    cout, taps, cin = w_khwc.shape
    wc = w_khwc.float().reshape(cout, kh, kw, cin).permute(0, 3, 1, 2)
    pad_h = kh // 2 if pad_h is None else pad_h
    pad_w = kw // 2 if pad_w is None else pad_w
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wc, None if bias is None else bias.float(), stride=stride, padding=(pad_h, pad_w))
    return y.permute(0, 2, 3, 1).to(BF16).contiguous()


def upsample2x_nhwc(x):
    return x.repeat_interleave(2, 1).repeat_interleave(2, 2).contiguous()


def softmax_rows(x, out=None):
    assert x.dtype == torch.float32
    return F.softmax(x, dim=-1).to(BF16)


def preprocess_frames(frames, rh, rw, top, left, oh, ow, mean, std, mode, flip=False, layout="image", dtype=torch.float32):
    """Transcription of preprocess.cu::preprocess_kernel (same formulas, fp64 accumulation), vectorised over pixels."""
    import numpy as np
    f = frames.cpu().numpy().astype(np.float64)
    n, h, w, _ = f.shape
    sy, sx = np.float32(h) / np.float32(rh), np.float32(w) / np.float32(rw)
    oy, ox = np.meshgrid(np.arange(oh), np.arange(ow), indexing="ij")
    ry = (oy + top).astype(np.float32)
    rx = ((ow - 1 - ox if flip else ox) + left).astype(np.float32)
    acc = np.zeros((n, oh, ow, 3))
    if mode == 0:
        fy, fx = np.maximum(sy * (ry + 0.5) - 0.5, 0), np.maximum(sx * (rx + 0.5) - 0.5, 0)
        y0, x0 = np.minimum(fy.astype(int), h - 1), np.minimum(fx.astype(int), w - 1)
        y1, x1 = np.minimum(y0 + 1, h - 1), np.minimum(x0 + 1, w - 1)
        ly, lx = (fy - y0)[None, ..., None], (fx - x0)[None, ..., None]
        acc = f[:, y0, x0] * (1 - ly) * (1 - lx) + f[:, y0, x1] * (1 - ly) * lx + f[:, y1, x0] * ly * (1 - lx) + f[:, y1, x1] * ly * lx
    elif mode == 1:
        A = -0.75

        def coeffs(t):
            x0, x1, x2, x3 = t + 1, t, 1 - t, 2 - t
            return [((A * x0 - 5 * A) * x0 + 8 * A) * x0 - 4 * A, ((A + 2) * x1 - (A + 3)) * x1 * x1 + 1,
                    ((A + 2) * x2 - (A + 3)) * x2 * x2 + 1, ((A * x3 - 5 * A) * x3 + 8 * A) * x3 - 4 * A]
        fy, fx = sy * (ry + 0.5) - 0.5, sx * (rx + 0.5) - 0.5
        yf, xf = np.floor(fy), np.floor(fx)
        cy, cx = coeffs(fy - yf), coeffs(fx - xf)
        for a in range(4):
            yy = np.clip(yf.astype(int) - 1 + a, 0, h - 1)
            for b in range(4):
                xx = np.clip(xf.astype(int) - 1 + b, 0, w - 1)
                acc += f[:, yy, xx] * (cy[a] * cx[b])[None, ..., None]
    else:
        def aa(x):
            a = -0.5
            x = np.abs(x)
            return np.where(x < 1, ((a + 2) * x - (a + 3)) * x * x + 1, np.where(x < 2, (((x - 5) * x + 8) * x - 4) * a, 0.0))

        def axis_weights(coord, scale, size):
            sc = max(float(scale), 1.0)
            sup = 2.0 * sc
            W = np.zeros((coord.shape[0], size))
            for i, c in enumerate(coord):
                ctr = float(scale) * (float(c) + 0.5)
                lo, hi = max(int(ctr - sup + 0.5), 0), min(int(ctr + sup + 0.5), size)
                ws = aa((np.arange(lo, hi) - ctr + 0.5) / sc)
                W[i, lo:hi] = ws / ws.sum()
            return W
        Wy = axis_weights(ry[:, 0], sy, h)            # [oh, h]
        Wx = axis_weights(rx[0, :], sx, w)            # [ow, w]
        acc = np.einsum("nyxc->nyxc", np.tensordot(np.tensordot(Wy, f, axes=([1], [1])), Wx, axes=([2], [1])).transpose(1, 0, 3, 2))
    v = (acc / 255.0 - np.asarray(mean)[None, None, None]) / np.asarray(std)[None, None, None]
    out = torch.from_numpy(v).permute(0, 3, 1, 2).to(dtype)
    return out.contiguous() if layout == "image" else out.permute(1, 0, 2, 3).contiguous()


def splice_multimodal(embed, feats, srcmap, out=None):
    src = srcmap.long().reshape(-1)
    v = torch.zeros((src.numel(), embed.shape[1]), dtype=BF16)
    tok = (src >= 0) & (src < embed.shape[0])
    v[tok] = embed[src[tok]]
    if feats is not None:                          # src = -(row + 1) selects feature row `row`; INT_MIN = padding (zeros)
        fr = (src < 0) & (src > -(2 ** 31)) & ((-(src + 1)) < feats.shape[0])
        v[fr] = feats[(-(src[fr] + 1))]
    v = v.reshape(*srcmap.shape, embed.shape[1])
    if out is not None:
        out.copy_(v.reshape(out.shape))
        return out
    return v


def add(a, b, out=None):
    v = (a.float().reshape(-1, b.numel()) + b.float().reshape(1, -1)).reshape(a.shape).to(BF16)
    if out is not None:
        out.copy_(v)
        return out
    return v


def patchify(pixels, patch, kpad):
    nb, c, h, w = pixels.shape
    cols = F.unfold(pixels.float(), patch, stride=patch).transpose(1, 2).reshape(-1, c * patch * patch)
    out = torch.zeros((cols.shape[0], kpad), dtype=BF16)
    out[:, :cols.shape[1]] = cols.to(BF16)
    return out


def vit_embed_ln(patch_out, cls, pos, ln_w, ln_b, nb, npatch, eps):
    d = patch_out.shape[-1]
    x = torch.cat([cls.float().view(1, 1, d).expand(nb, 1, d), patch_out.float().view(nb, npatch, d)], 1) + pos.float()[None]
    return F.layer_norm(x, (d,), ln_w.float(), ln_b.float(), eps).to(BF16)


def seem_attn_mask(mask_logits, h2, w2):
    """fp32 [Q, H, W] -> uint8 [Q, h2*w2]: bilinear resize, masked where sigmoid < 0.5 (logit < 0), fully masked rows reset."""
    r = F.interpolate(mask_logits[None].float(), size=(h2, w2), mode="bilinear", align_corners=False)[0].flatten(1)
    m = r < 0
    m[m.all(dim=1)] = False
    return m.to(torch.uint8)


def resize_bilinear_nhwc(x, h2, w2):
    """bf16 NHWC [nb, H, W, C] -> [nb, h2, w2, C], bilinear, align_corners=False."""
    r = F.interpolate(x.float().permute(0, 3, 1, 2), size=(h2, w2), mode="bilinear", align_corners=False)
    return r.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()


def attention_short(q, k, v, scale=None, out=None):
    """q/k/v [nseq, S, H, D] or [outer, inner, S, H, D] strided views; attention over S per (sequence, head)."""
    D = q.shape[-1]
    scale = D ** -0.5 if scale is None else scale
    s = torch.einsum("...qhd,...khd->...hqk", q.float(), k.float()) * scale
    o = torch.einsum("...hqk,...khd->...qhd", s.softmax(-1), v.float()).to(BF16)
    if out is not None:
        out.copy_(o)
        return out
    return o.contiguous()


def cfg_combine(y, u, scale):
    return (u + scale * (y - u)).contiguous()


def rope_kv_append(qkv, positions, n_heads, head_dim, theta, k_pages=None, v_pages=None, block_table=None,
                   batch_of_token=None, slot_of_token=None, page_size=0):
    """rotate_half RoPE on the q | k thirds of every row in place (the K/V page scatter only matters for decode)."""
    T = qkv.shape[0]
    x = qkv.float().view(T, 3, n_heads, head_dim)
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    fr = positions.float()[:, None] * inv
    emb = torch.cat([fr, fr], -1)
    cos, sin = emb.cos()[:, None, :], emb.sin()[:, None, :]
    rot = lambda t: torch.cat([-t[..., head_dim // 2:], t[..., :head_dim // 2]], -1)
    for j in (0, 1):
        x[:, j] = x[:, j] * cos + rot(x[:, j]) * sin
    qkv.copy_(x.view(T, -1).to(BF16))
    if k_pages is not None and block_table is not None:   # K/V scatter into the paged cache [pages, H, page_size, D]
        xb = qkv.view(T, 3, n_heads, head_dim)
        for t in range(T):
            slot = int(slot_of_token[t]) if slot_of_token is not None else int(positions[t])
            if slot < 0:
                continue
            b = int(batch_of_token[t]) if batch_of_token is not None else 0
            page = int(block_table[b, slot // page_size])
            k_pages[page, :, slot % page_size] = xb[t, 1]
            v_pages[page, :, slot % page_size] = xb[t, 2]
    return qkv


def row_rstd(x, eps):
    x2 = x.float().reshape(-1, x.shape[-1])
    return torch.rsqrt(x2.pow(2).mean(-1) + eps)


def rope_table(positions, head_dim, theta, out=None):
    """fp32 [B, head_dim] = cos | sin of inv_freq_i * position (vb200_rope_table)."""
    half = head_dim // 2
    inv = torch.exp2(-(torch.arange(half, dtype=torch.float32) / half) * math.log2(theta))
    ang = positions.float()[:, None] * inv[None]
    tab = torch.cat([ang.cos(), ang.sin()], -1)
    if out is not None:
        out.copy_(tab)
        return out
    return tab


def attn_decode_rope(qkv, table, k_pages, v_pages, block_table, kv_len, n_heads, head_dim, page_size, max_kv_len, scale=None,
                     out=None):
    """One decode token per sequence: RoPE on q / k from the cos|sin table, K/V of the new token appended at slot
    kv_len - 1 of the paged cache, softmax attention over the kv_len keys (vb200_attn_decode_rope)."""
    B = qkv.shape[0]
    half = head_dim // 2
    scale = head_dim ** -0.5 if scale is None else scale
    x = qkv[:, :3 * n_heads * head_dim].float().view(B, 3, n_heads, head_dim)
    cos = torch.cat([table[:, :half], table[:, :half]], -1)[:, None]
    sin = torch.cat([table[:, half:], table[:, half:]], -1)[:, None]
    rot = lambda t: torch.cat([-t[..., half:], t[..., :half]], -1)
    q = (x[:, 0] * cos + rot(x[:, 0]) * sin).to(BF16).float()
    k = (x[:, 1] * cos + rot(x[:, 1]) * sin).to(BF16)
    v = x[:, 2].to(BF16)
    res = torch.empty((B, n_heads * head_dim), dtype=BF16) if out is None else out
    for b in range(B):
        n = int(kv_len[b])
        slot = n - 1
        page = int(block_table[b, slot // page_size])
        k_pages[page, :, slot % page_size] = k[b]
        v_pages[page, :, slot % page_size] = v[b]
        pages = block_table[b, :(n + page_size - 1) // page_size].long()
        kk = k_pages[pages].permute(1, 0, 2, 3).reshape(n_heads, -1, head_dim)[:, :n].float()
        vv = v_pages[pages].permute(1, 0, 2, 3).reshape(n_heads, -1, head_dim)[:, :n].float()
        p = torch.softmax(torch.einsum("hd,hnd->hn", q[b], kk) * scale, -1)
        res[b] = torch.einsum("hn,hnd->hd", p, vv).reshape(-1).to(BF16)
    return res


def argmax_rows(logits, out=None):
    idx = logits.float().reshape(-1, logits.shape[-1]).argmax(-1)
    if out is not None:
        out.copy_(idx)
        return out
    return idx


def argmax_advance(logits, out_idx, next_src=None, positions=None, kv_len=None, token_log=None, prompt_len=None):
    """vb200_argmax_advance: arg-max + the device-side decode bookkeeping."""
    idx = logits.float().argmax(-1)
    out_idx.copy_(idx)
    if next_src is not None:
        next_src.copy_(idx.to(next_src.dtype))
    if token_log is not None:
        for b in range(idx.shape[0]):
            step = int(kv_len[b]) - int(prompt_len[b])
            if 0 <= step < token_log.shape[1]:
                token_log[b, step] = idx[b]
    if positions is not None:
        positions.add_(1)
    if kv_len is not None:
        kv_len.add_(1)
    return out_idx


def region_mask_pool(feats, boxes, image_size):
    """feats [B, g*g, C], boxes fp32 [B, 4] -> [B, C]: the reference's MaskPooling incl. the x-indexes-rows quirk
    (region_extractor/layer.py:27-43, 77-112)."""
    B, n, c = feats.shape
    g = int(n ** 0.5)
    m = torch.zeros((B, 1, image_size, image_size))
    for b in range(B):
        x1, y1, x2, y2 = [float(v) for v in boxes[b]]
        m[b, 0, int(x1):int(x2), int(y1):int(y2)] = 1
    m = (F.interpolate(m, size=(g, g), mode="bilinear", align_corners=False) > 0).float()
    den = m.sum(dim=(-1, -2), keepdim=True) + 1e-8
    f = feats.float().reshape(B, g, g, c).permute(0, 3, 1, 2)
    return torch.einsum("bchw,bqhw->bqc", f, m / den).reshape(B, c).to(BF16)


def add_rowgroup(x, table, group_rows, period, out=None):
    rows = x.reshape(-1, x.shape[-1])
    idx = (torch.arange(rows.shape[0]) // group_rows) % period
    v = (rows.float() + table.float()[idx]).to(BF16).reshape(x.shape)
    if out is not None:
        out.copy_(v)
        return out
    return v


def install(monkeypatch):
    """Replace the kernel-launching entry points of vitron_b200.ops with the statements above."""
    from vitron_b200 import ops
    for name in ("gemm", "layernorm", "layernorm_add", "pack_dwconv_weight", "dwconv_nhwc", "colmean", "focal_modulate",
                 "mul_rows", "im2col_nchw", "pack_conv_weight", "conv_nhwc", "groupnorm_nhwc", "conv_nhwc_direct",
                 "upsample2x_nhwc", "softmax_rows", "preprocess_frames", "pack_glu_weight", "attention",
                 "splice_multimodal", "add", "patchify", "vit_embed_ln", "seem_attn_mask", "resize_bilinear_nhwc", "attention_short",
                 "cfg_combine", "rope_kv_append", "region_mask_pool", "add_rowgroup", "row_rstd", "rope_table", "attn_decode_rope",
                 "argmax_rows", "argmax_advance"):
        monkeypatch.setattr(ops, name, globals()[name])
