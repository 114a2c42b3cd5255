"""Parameter names / shapes of the reference modules on the path, used to
random-initialise full-size models on the device for benchmarks and to check state-dict coverage."""
import torch

BF16 = torch.bfloat16


def llama_shapes(c):
    d, f, V = c.hidden_size, c.intermediate_size, c.vocab_size
    s = {"model.embed_tokens.weight": [V, d], "lm_head.weight": [V, d], "model.norm.weight": [d]}
    for i in range(c.num_hidden_layers):
        p = f"model.layers.{i}."
        for n in "qkvo":
            s[p + f"self_attn.{n}_proj.weight"] = [d, d]
        s[p + "mlp.gate_proj.weight"] = [f, d]
        s[p + "mlp.up_proj.weight"] = [f, d]
        s[p + "mlp.down_proj.weight"] = [d, f]
        s[p + "input_layernorm.weight"] = [d]
        s[p + "post_attention_layernorm.weight"] = [d]
    return s


def vit_shapes(c, prefix):
    d, f = c.hidden_size, c.intermediate_size
    npatch = (c.image_size // c.patch_size) ** 2
    s = {prefix + "embeddings.class_embedding": [d],
         prefix + "embeddings.patch_embedding.weight": [d, c.num_channels, c.patch_size, c.patch_size],
         prefix + "embeddings.position_embedding.weight": [npatch + 1, d],
         prefix + "pre_layrnorm.weight": [d], prefix + "pre_layrnorm.bias": [d]}
    for i in range(c.num_hidden_layers):
        p = prefix + f"encoder.layers.{i}."
        attns = ["self_attn"] + (["temporal_attn"] if c.add_time_attn else [])
        for a in attns:
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                s[p + f"{a}.{n}.weight"] = [d, d]
                s[p + f"{a}.{n}.bias"] = [d]
        norms = ["layer_norm1", "layer_norm2"] + (["temporal_layer_norm1"] if c.add_time_attn else [])
        for n in norms:
            s[p + n + ".weight"] = [d]
            s[p + n + ".bias"] = [d]
        if c.add_time_attn:
            s[p + "temporal_embedding"] = [1, c.num_frames, d]
        s[p + "mlp.fc1.weight"] = [f, d]
        s[p + "mlp.fc1.bias"] = [f]
        s[p + "mlp.fc2.weight"] = [d, f]
        s[p + "mlp.fc2.bias"] = [d]
    return s


def projector_shapes(mm_hidden, hidden, depth=2, prefix="model.mm_projector."):
    s = {}
    for i in range(depth):
        s[prefix + f"{2 * i}.weight"] = [hidden, mm_hidden if i == 0 else hidden]
        s[prefix + f"{2 * i}.bias"] = [hidden]
    return s


def region_shapes(in_dim, out_dim, prefix="model.region_extractor."):
    s = {}
    for i, (o, k) in enumerate([(out_dim, in_dim), (out_dim, out_dim), (out_dim, out_dim)]):
        s[prefix + f"region_linear.layers.{i}.weight"] = [o, k]
        s[prefix + f"region_linear.layers.{i}.bias"] = [o]
    s[prefix + "loc_encoder.loc_encoder.0.weight"] = [out_dim // 2, 4]
    s[prefix + "loc_encoder.loc_encoder.0.bias"] = [out_dim // 2]
    s[prefix + "loc_encoder.loc_encoder.2.weight"] = [out_dim, out_dim // 2]
    s[prefix + "loc_encoder.loc_encoder.2.bias"] = [out_dim]
    return s


def vitron_shapes(cfg):
    s = llama_shapes(cfg.llm)
    if cfg.vision is not None:
        s.update(vit_shapes(cfg.vision, "model.image_tower.image_tower."))
    if cfg.video is not None:
        s.update(vit_shapes(cfg.video, "model.video_tower.video_tower."))
    vc = cfg.vision or cfg.video
    if vc is not None:
        s.update(projector_shapes(vc.hidden_size, cfg.llm.hidden_size))
        s.update(region_shapes(vc.hidden_size, cfg.llm.hidden_size))
    return s


def unet_shapes(cfg):
    """Reference parameter names / shapes of UNetSD_I2VGen for a config (block plan = the drop-in's own,
    which mirrors unet_i2vgen.py:133-233)."""
    from .unet_i2vgen import UNetSD_I2VGen
    dim, ed, ctx = cfg["dim"], cfg["dim"] * 4, cfg["context_dim"]
    cd = cfg["in_dim"]
    s = {}

    def lin(p, o, i, bias=True):
        s[p + ".weight"] = [o, i]
        if bias:
            s[p + ".bias"] = [o]

    def norm(p, c):
        s[p + ".weight"] = [c]
        s[p + ".bias"] = [c]

    def tblock(p, inner, cdim):
        for a, kd in (("attn1", inner), ("attn2", cdim)):
            lin(f"{p}.{a}.to_q", inner, inner, False); lin(f"{p}.{a}.to_k", inner, kd, False); lin(f"{p}.{a}.to_v", inner, kd, False)
            lin(f"{p}.{a}.to_out.0", inner, inner)
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{p}.{n}", inner)
        lin(f"{p}.ff.net.0.proj", inner * 8, inner); lin(f"{p}.ff.net.2", inner, inner * 4)

    for n, i in (("time_embed", dim), ("fps_embedding", dim), ("context_embedding", cfg["y_dim"])):
        lin(n + ".0", ed, i); lin(n + ".2", ctx * cfg["num_tokens"] if n == "context_embedding" else ed, ed)
    for k, (o, i) in zip((0, 2, 4), ((cd * 4, 4), (cd * 4, cd * 4), (cd, cd * 4))):
        s[f"local_image_concat.{k}.weight"] = [o, i, 3, 3]; s[f"local_image_concat.{k}.bias"] = [o]
    p = "local_temporal_encoder.layers.0."
    norm(p + "0.norm", cd); lin(p + "0.fn.to_qkv", cd * 2 * 3, cd, False); lin(p + "0.fn.to_out.0", cd, cd * 2)
    lin(p + "1.net.0.0", cd * 4, cd); lin(p + "1.net.2", cd, cd * 4)
    for k, (o, i) in zip((0, 3, 5), ((cd * 8, 4), (cd * 16, cd * 8), (1024, cd * 16))):
        s[f"local_image_embedding.{k}.weight"] = [o, i, 3, 3]; s[f"local_image_embedding.{k}.bias"] = [o]
    m = UNetSD_I2VGen(**cfg, device="cpu")
    blocks = list(m._all_blocks())
    for kind, p, ci, co in blocks:
        if kind == "conv_in":
            s[p + ".weight"] = [dim, cfg["in_dim"] + cd, 3, 3]; s[p + ".bias"] = [dim]
        elif kind == "res":
            norm(p + ".in_layers.0", ci); s[p + ".in_layers.2.weight"] = [co, ci, 3, 3]; s[p + ".in_layers.2.bias"] = [co]
            lin(p + ".emb_layers.1", co, ed); norm(p + ".out_layers.0", co)
            s[p + ".out_layers.3.weight"] = [co, co, 3, 3]; s[p + ".out_layers.3.bias"] = [co]
            if ci != co:
                s[p + ".skip_connection.weight"] = [co, ci, 1, 1]; s[p + ".skip_connection.bias"] = [co]
            for c_, l_ in ((1, 2), (2, 3), (3, 3), (4, 3)):
                norm(f"{p}.temopral_conv.conv{c_}.0", co)
                s[f"{p}.temopral_conv.conv{c_}.{l_}.weight"] = [co, co, 3, 1, 1]; s[f"{p}.temopral_conv.conv{c_}.{l_}.bias"] = [co]
        elif kind == "st":
            norm(p + ".norm", ci); lin(p + ".proj_in", ci, ci); lin(p + ".proj_out", ci, ci)
            tblock(p + ".transformer_blocks.0", ci, ctx)
        elif kind == "tt":
            inner = co * cfg["head_dim"]
            norm(p + ".norm", ci)
            s[p + ".proj_in.weight"] = [inner, ci, 1]; s[p + ".proj_in.bias"] = [inner]
            s[p + ".proj_out.weight"] = [ci, inner, 1]; s[p + ".proj_out.bias"] = [ci]
            tblock(p + ".transformer_blocks.0", inner, inner)
        elif kind == "down":
            s[p + ".op.weight"] = [co, ci, 3, 3]; s[p + ".op.bias"] = [co]
        elif kind == "up":
            s[p + ".conv.weight"] = [co, ci, 3, 3]; s[p + ".conv.bias"] = [co]
    fd = dim
    norm("out.0", fd); s["out.2.weight"] = [cfg["out_dim"], fd, 3, 3]; s["out.2.bias"] = [cfg["out_dim"]]
    return s


def seem_shapes(in_channels=(192, 384, 768, 1536), conv_dim=512, ffn=2048, queries=101, enc_layers=6, dec_layers=9,
                dim_proj=512):
    """Parameter names / shapes of SEEM's sem_seg_head.{pixel_decoder,predictor} that the seg path reads
    (transformer_encoder_fpn.py:23-308, seem.py:193-392)."""
    C, s = conv_dim, {}

    def lin(p, o, i):
        s[p + ".weight"], s[p + ".bias"] = [o, i], [o]

    def norm(p):
        s[p + ".weight"], s[p + ".bias"] = [C], [C]

    def mha(p):
        s[p + ".in_proj_weight"], s[p + ".in_proj_bias"] = [3 * C, C], [3 * C]
        lin(p + ".out_proj", C, C)

    pd = "pixel_decoder."
    s[pd + "input_proj.weight"], s[pd + "input_proj.bias"] = [C, in_channels[-1], 1, 1], [C]
    for i in range(enc_layers):
        q = pd + f"transformer.encoder.layers.{i}"
        mha(q + ".self_attn"); lin(q + ".linear1", ffn, C); lin(q + ".linear2", C, ffn); norm(q + ".norm1"); norm(q + ".norm2")
    for idx, cin in enumerate(in_channels):
        s[pd + f"layer_{idx + 1}.weight"] = [C, C, 3, 3]
        norm(pd + f"layer_{idx + 1}.norm")
        if idx != len(in_channels) - 1:
            s[pd + f"adapter_{idx + 1}.weight"] = [C, cin, 1, 1]
            norm(pd + f"adapter_{idx + 1}.norm")
    s[pd + "mask_features.weight"], s[pd + "mask_features.bias"] = [C, C, 3, 3], [C]
    pr = "predictor."
    for i in range(dec_layers):
        mha(pr + f"transformer_cross_attention_layers.{i}.multihead_attn"); norm(pr + f"transformer_cross_attention_layers.{i}.norm")
        mha(pr + f"transformer_self_attention_layers.{i}.self_attn"); norm(pr + f"transformer_self_attention_layers.{i}.norm")
        lin(pr + f"transformer_ffn_layers.{i}.linear1", ffn, C); lin(pr + f"transformer_ffn_layers.{i}.linear2", C, ffn)
        norm(pr + f"transformer_ffn_layers.{i}.norm")
    norm(pr + "decoder_norm")
    for n, shape in (("query_feat.weight", [queries, C]), ("query_embed.weight", [queries, C]), ("level_embed.weight", [3, C])):
        s[pr + n] = shape
    for i in range(3):
        lin(pr + f"mask_embed.layers.{i}", C, C)
    s[pr + "class_embed"] = [C, dim_proj]
    return s


def focalnet_shapes(cfg):
    """Reference FocalNet parameter names / shapes (backbone/focal.py:340-437); cfg keys as oracle.restate_focal.FOCAL_L
    (embed_dim, depths, focal_levels, focal_windows, mlp_ratio, use_conv_embed, use_postln_in_modulation,
    use_layerscale, patch_norm, out_indices)."""
    s = {}
    E = cfg["embed_dim"]
    k = 7 if cfg["use_conv_embed"] else cfg["patch_size"]
    s["patch_embed.proj.weight"], s["patch_embed.proj.bias"] = [E, 3, k, k], [E]
    if cfg["patch_norm"]:
        s["patch_embed.norm.weight"], s["patch_embed.norm.bias"] = [E], [E]
    n = len(cfg["depths"])
    for i in range(n):
        C = E * 2 ** i
        L, win = cfg["focal_levels"][i], cfg["focal_windows"][i]
        hid = int(C * cfg["mlp_ratio"])
        for j in range(cfg["depths"][i]):
            p = f"layers.{i}.blocks.{j}."
            for nm in ("norm1", "norm2"):
                s[p + nm + ".weight"], s[p + nm + ".bias"] = [C], [C]
            s[p + "modulation.f.weight"], s[p + "modulation.f.bias"] = [2 * C + L + 1, C], [2 * C + L + 1]
            s[p + "modulation.h.weight"], s[p + "modulation.h.bias"] = [C, C, 1, 1], [C]
            s[p + "modulation.proj.weight"], s[p + "modulation.proj.bias"] = [C, C], [C]
            if cfg["use_postln_in_modulation"]:
                s[p + "modulation.ln.weight"], s[p + "modulation.ln.bias"] = [C], [C]
            for l in range(L):
                kk = 2 * l + win
                s[p + f"modulation.focal_layers.{l}.0.weight"] = [C, 1, kk, kk]
            s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"] = [hid, C], [hid]
            s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"] = [C, hid], [C]
            if cfg["use_layerscale"]:
                s[p + "gamma_1"], s[p + "gamma_2"] = [C], [C]
        if i < n - 1:
            p = f"layers.{i}.downsample."
            kd = 3 if cfg["use_conv_embed"] else 2
            s[p + "proj.weight"], s[p + "proj.bias"] = [2 * C, C, kd, kd], [2 * C]
            s[p + "norm.weight"], s[p + "norm.bias"] = [2 * C], [2 * C]
        if i in cfg["out_indices"]:
            s[f"norm{i}.weight"], s[f"norm{i}.bias"] = [C], [C]
    return s


def vae_shapes(dd, embed_dim=4):
    """Reference AutoencoderKL parameter names / shapes (i2vgen-xl tools/modules/autoencoder.py:31-62, 483-651)."""
    s = {}
    ch, mult, nrb, z = dd["ch"], tuple(dd["ch_mult"]), dd["num_res_blocks"], dd["z_channels"]

    def conv(p, cout, cin, k):
        s[p + "weight"], s[p + "bias"] = [cout, cin, k, k], [cout]

    def norm(p, c):
        s[p + "weight"], s[p + "bias"] = [c], [c]

    def res(p, cin, cout):
        norm(p + "norm1.", cin); conv(p + "conv1.", cout, cin, 3); norm(p + "norm2.", cout); conv(p + "conv2.", cout, cout, 3)
        if cin != cout:
            conv(p + "nin_shortcut.", cout, cin, 1)

    def attn(p, c):
        norm(p + "norm.", c)
        for n in ("q.", "k.", "v.", "proj_out."):
            conv(p + n, c, c, 1)

    conv("encoder.conv_in.", ch, dd["in_channels"], 3)
    in_mult = (1,) + mult
    block_in = ch
    for i in range(len(mult)):
        block_in, block_out = ch * in_mult[i], ch * mult[i]
        for j in range(nrb):
            res(f"encoder.down.{i}.block.{j}.", block_in, block_out)
            block_in = block_out
        if i != len(mult) - 1:
            conv(f"encoder.down.{i}.downsample.conv.", block_in, block_in, 3)
    res("encoder.mid.block_1.", block_in, block_in); attn("encoder.mid.attn_1.", block_in); res("encoder.mid.block_2.", block_in, block_in)
    norm("encoder.norm_out.", block_in)
    conv("encoder.conv_out.", 2 * z if dd["double_z"] else z, block_in, 3)
    block_in = ch * mult[-1]
    conv("decoder.conv_in.", block_in, z, 3)
    res("decoder.mid.block_1.", block_in, block_in); attn("decoder.mid.attn_1.", block_in); res("decoder.mid.block_2.", block_in, block_in)
    for i in reversed(range(len(mult))):
        block_out = ch * mult[i]
        for j in range(nrb + 1):
            res(f"decoder.up.{i}.block.{j}.", block_in, block_out)
            block_in = block_out
        if i != 0:
            conv(f"decoder.up.{i}.upsample.conv.", block_in, block_in, 3)
    norm("decoder.norm_out.", block_in)
    conv("decoder.conv_out.", dd["out_ch"], block_in, 3)
    conv("quant_conv.", 2 * embed_dim, 2 * z, 1)
    conv("post_quant_conv.", z, embed_dim, 1)
    return s


def gligen_unet_shapes(cfg):
    """Reference GLIGEN UNetModel parameter names / shapes (openaimodel.py:234-386, attention.py:285-386,
    positionnet.py:9-28) for fuser_type='gatedSA'."""
    s = {}
    mc, mult, nrb = cfg["model_channels"], tuple(cfg["channel_mult"]), cfg["num_res_blocks"]
    att, heads, depth = list(cfg["attention_resolutions"]), cfg.get("num_heads", 8), cfg.get("transformer_depth", 1)
    ctx, pos_len = cfg["context_dim"], cfg.get("positive_len", 768)
    ted = 4 * mc

    def lin(p, o, i, bias=True):
        s[p + "weight"] = [o, i]
        if bias:
            s[p + "bias"] = [o]

    def conv(p, o, i, k):
        s[p + "weight"], s[p + "bias"] = [o, i, k, k], [o]

    def norm(p, c):
        s[p + "weight"], s[p + "bias"] = [c], [c]

    def res(p, cin, cout):
        norm(p + "in_layers.0.", cin); conv(p + "in_layers.2.", cout, cin, 3); lin(p + "emb_layers.1.", cout, ted)
        norm(p + "out_layers.0.", cout); conv(p + "out_layers.3.", cout, cout, 3)
        if cin != cout:
            conv(p + "skip_connection.", cout, cin, 1)

    def attn(p, q, kdim, self_attn):
        lin(p + "to_q.", q, q, False); lin(p + "to_k.", q, q if self_attn else kdim, False); lin(p + "to_v.", q, q if self_attn else kdim, False)
        lin(p + "to_out.0.", q, q)

    def ff(p, d):
        lin(p + "net.0.proj.", 8 * d, d); lin(p + "net.2.", d, 4 * d)

    def st(p, c):
        norm(p + "norm.", c); conv(p + "proj_in.", c, c, 1); conv(p + "proj_out.", c, c, 1)
        for d in range(depth):
            b = p + f"transformer_blocks.{d}."
            attn(b + "attn1.", c, c, True); ff(b + "ff.", c); attn(b + "attn2.", c, ctx, False)
            for i in (1, 2, 3):
                norm(b + f"norm{i}.", c)
            f = b + "fuser."
            lin(f + "linear.", c, ctx); attn(f + "attn.", c, c, True); ff(f + "ff.", c)
            norm(f + "norm1.", c); norm(f + "norm2.", c)
            s[f + "alpha_attn"], s[f + "alpha_dense"] = [], []

    lin("time_embed.0.", ted, mc); lin("time_embed.2.", ted, ted)
    tin = 2 * cfg["in_channels"] + 1 if cfg.get("is_inpaint", False) else cfg["in_channels"]
    conv("input_blocks.0.0.", mc, tin, 3)
    chans, ch, ds, idx = [mc], mc, 1, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            res(f"input_blocks.{idx}.0.", ch, m * mc)
            ch = m * mc
            if ds in att:
                st(f"input_blocks.{idx}.1.", ch)
            chans.append(ch)
            idx += 1
        if level != len(mult) - 1:
            conv(f"input_blocks.{idx}.0.op.", ch, ch, 3)
            chans.append(ch)
            ds *= 2
            idx += 1
    res("middle_block.0.", ch, ch); st("middle_block.1.", ch); res("middle_block.2.", ch, ch)
    idx = 0
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            res(f"output_blocks.{idx}.0.", ch + chans.pop(), mc * m)
            ch = mc * m
            j = 1
            if ds in att:
                st(f"output_blocks.{idx}.{j}.", ch)
                j += 1
            if level and i == nrb:
                conv(f"output_blocks.{idx}.{j}.conv.", ch, ch, 3)
                ds //= 2
            idx += 1
    norm("out.0.", ch); conv("out.2.", cfg["out_channels"], mc, 3)
    lin("position_net.linears.0.", 512, pos_len + 64); lin("position_net.linears.2.", 512, 512); lin("position_net.linears.4.", ctx, 512)
    s["position_net.null_positive_feature"], s["position_net.null_position_feature"] = [pos_len], [64]
    return s


def openclip_shapes(cfg):
    """open_clip CLIP state-dict names / shapes under the embedder's `model.` attribute (cfg as clip_embedder.VIT_H_14)."""
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


def random_state_dict(shapes, device, seed=0, std=0.02):
    """N(0, std) weights, unit norm gains, zero biases, generated on `device`."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for name, shape in shapes.items():
        low = name.lower()
        if len(shape) == 1 and low.endswith("weight"):  # every 1-D ".weight" on the path is a norm gain
            out[name] = torch.ones(shape, dtype=BF16, device=device)
        elif len(shape) == 0:
            out[name] = torch.zeros(shape, dtype=BF16, device=device)
        elif low.endswith("bias"):
            out[name] = torch.zeros(shape, dtype=BF16, device=device)
        else:
            out[name] = (torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std).to(BF16)
    return out
