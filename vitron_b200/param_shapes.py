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


def random_state_dict(shapes, device, seed=0, std=0.02):
    """N(0, std) weights, unit norm gains, zero biases, generated on `device`."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for name, shape in shapes.items():
        low = name.lower()
        if len(shape) == 1 and low.endswith("weight") and ("norm" in low):
            out[name] = torch.ones(shape, dtype=BF16, device=device)
        elif low.endswith("bias"):
            out[name] = torch.zeros(shape, dtype=BF16, device=device)
        else:
            out[name] = (torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std).to(BF16)
    return out
