"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of what i2vgen-xl's FrozenOpenCLIP*Embedder classes compute
(modules/i2vgen-xl/tools/modules/clip_embedder.py:41-66, 179-214); never imported by the product package.

The reference's own code (clip_embedder.py): token_embedding + positional_embedding, resblocks[:len - layer_idx] with
the text tower's causal `attn_mask`, `ln_final`, `x[arange, text.argmax(-1)] @ text_projection`, `model.encode_image`.
The transformer arithmetic is THIRD-PARTY: `open_clip` (imported at clip_embedder.py:4; version unpinned in
modules/i2vgen-xl/requirements.txt: `open-clip-torch`), absent from this image. Its published algorithm
(open_clip/transformer.py) is restated here: ResidualAttentionBlock x = x + attn(ln_1(x)), x = x + mlp(ln_2(x)) with
nn.MultiheadAttention (in_proj_weight [3d, d]) and mlp = c_fc -> GELU -> c_proj; VisionTransformer = conv1 (no bias) ->
[class_embedding ; patches] + positional_embedding -> ln_pre -> resblocks -> ln_post(x[:, 0]) @ proj.
Parity status: UNPINNED against open_clip itself (not installable here, no vectors in the reference);
cross-checked against the independent implementation in `transformers` (CLIPTextModel / CLIPVisionModel with the same
weights mapped, tests/test_oracle_cpu.py::test_openclip_restatement_matches_transformers_clip).
"""
import torch
import torch.nn.functional as F


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"].float(), sd[p + "bias"].float(), 1e-5)


def resblock(x, sd, p, heads, causal):
    """x [B, S, d]."""
    B, S, d = x.shape
    hd = d // heads
    h = _ln(x, sd, p + "ln_1.")
    qkv = F.linear(h, sd[p + "attn.in_proj_weight"].float(), sd[p + "attn.in_proj_bias"].float()).view(B, S, 3, heads, hd)
    q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) * hd ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(S, S, dtype=torch.bool).triu(1), float("-inf"))
    a = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, S, d)
    x = x + F.linear(a, sd[p + "attn.out_proj.weight"].float(), sd[p + "attn.out_proj.bias"].float())
    h = F.gelu(F.linear(_ln(x, sd, p + "ln_2."), sd[p + "mlp.c_fc.weight"].float(), sd[p + "mlp.c_fc.bias"].float()))
    return x + F.linear(h, sd[p + "mlp.c_proj.weight"].float(), sd[p + "mlp.c_proj.bias"].float())


def encode_text(sd, tokens, cfg, layer_idx, p="model."):
    """-> (xt [B, embed], x [B, S, width]) as FrozenOpenCLIPTtxtVisualEmbedder.encode_with_transformer."""
    t = cfg["text"]
    x = sd[p + "token_embedding.weight"].float()[tokens] + sd[p + "positional_embedding"].float()[:tokens.shape[1]]
    for i in range(t["layers"] - layer_idx):
        x = resblock(x, sd, p + f"transformer.resblocks.{i}.", t["heads"], True)
    x = _ln(x, sd, p + "ln_final.")
    xt = x[torch.arange(x.shape[0]), tokens.argmax(dim=-1)] @ sd[p + "text_projection"].float()
    return xt, x


def encode_image(sd, image, cfg, p="model.visual."):
    v = cfg["vision"]
    x = F.conv2d(image.float(), sd[p + "conv1.weight"].float(), None, stride=v["patch_size"])
    x = x.flatten(2).transpose(1, 2)
    cls = sd[p + "class_embedding"].float().view(1, 1, -1).expand(x.shape[0], -1, -1)
    x = torch.cat([cls, x], dim=1) + sd[p + "positional_embedding"].float()
    x = _ln(x, sd, p + "ln_pre.")
    for i in range(v["layers"]):
        x = resblock(x, sd, p + f"transformer.resblocks.{i}.", v["heads"], False)
    return _ln(x[:, 0], sd, p + "ln_post.") @ sd[p + "proj"].float()


def openclip_shapes(cfg):
    """open_clip CLIP state-dict names / shapes under the embedder's `model.` attribute."""
    s = {}
    t, v, e = cfg["text"], cfg["vision"], cfg["embed_dim"]

    def blocks(p, n, d, mlp):
        for i in range(n):
            q = p + f"resblocks.{i}."
            s[q + "attn.in_proj_weight"], s[q + "attn.in_proj_bias"] = [3 * d, d], [3 * d]
            s[q + "attn.out_proj.weight"], s[q + "attn.out_proj.bias"] = [d, d], [d]
            for n_ in ("ln_1.", "ln_2."):
                s[q + n_ + "weight"], s[q + n_ + "bias"] = [d], [d]
            s[q + "mlp.c_fc.weight"], s[q + "mlp.c_fc.bias"] = [mlp, d], [mlp]
            s[q + "mlp.c_proj.weight"], s[q + "mlp.c_proj.bias"] = [d, mlp], [d]

    d = t["width"]
    s["model.token_embedding.weight"], s["model.positional_embedding"] = [t["vocab_size"], d], [t["context_length"], d]
    blocks("model.transformer.", t["layers"], d, 4 * d)
    s["model.ln_final.weight"], s["model.ln_final.bias"], s["model.text_projection"] = [d], [d], [d, e]
    d = v["width"]
    npatch = (v["image_size"] // v["patch_size"]) ** 2
    s["model.visual.conv1.weight"] = [d, 3, v["patch_size"], v["patch_size"]]
    s["model.visual.class_embedding"], s["model.visual.positional_embedding"] = [d], [npatch + 1, d]
    for n_ in ("ln_pre.", "ln_post."):
        s["model.visual." + n_ + "weight"], s["model.visual." + n_ + "bias"] = [d], [d]
    blocks("model.visual.transformer.", v["layers"], d, v["mlp"])
    s["model.visual.proj"] = [d, e]
    return s
