"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference's vision-LLM forward path.

Plain torch on the CPU, no kernels from vitron_b200, never imported by the product package
(only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it).  Every function
follows the reference (or, where the arithmetic lives in the pinned third-party dependency
transformers==4.31.0 — pyproject.toml:16 — its published algorithm as called by the reference):

  llama_*            transformers 4.31 modeling_llama (LlamaRMSNorm / rotate_half RoPE / eager
                     softmax attention / SiLU-gated MLP); in-tree restatement of the attention
                     sequence: vitron/train/llama_flash_attn_monkey_patch.py:30-66; call site
                     vitron/model/language_model/llava_llama.py:91-102
  clip_vit_hidden    languagebind/image/modeling_image.py:86-158,610-672 and
                     video/modeling_video.py:86-158,610-676 (+ HF CLIPVisionEmbeddings /
                     CLIPAttention / CLIPMLP); tower select: languagebind/__init__.py:96-121,182-204
  projector          vitron/model/multimodal_projector/builder.py:33-51
  region_extractor   vitron/model/region_extractor/layer.py:27-43,77-130
  splice             vitron/model/llava_arch.py:189-573
  greedy_generate    HF greedy search as driven by llava_llama.py:57-114 (cache-less recompute,
                     equal to the cached path — SURVEY.md Appendix C item 10)

Parity status: PINNED — tests/test_oracle_cpu.py checks these functions against the unmodified
reference imported through oracle/refshim.py (when /root/reference exists) and against the
committed golden vectors in tests/golden/ that oracle/gen_golden.py produced from that reference.
State dicts use the reference's parameter names.
"""
import math

import torch
import torch.nn.functional as F

IGNORE_INDEX, IMAGE_TOKEN_INDEX, OBJS_TOKEN_INDEX = -100, -200, -300


# ------------------------------------------------------------------ LLaMA
def rms_norm(x, w, eps):
    v = x.float().pow(2).mean(-1, keepdim=True)
    return w.float() * (x.float() * torch.rsqrt(v + eps))


def rope_cos_sin(positions, dim, theta):
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    freqs = positions.float()[..., None] * inv
    emb = torch.cat([freqs, freqs], -1)
    return emb.cos(), emb.sin()


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], -1)


def llama_forward(sd, cfg, inputs_embeds, lens=None):
    """inputs_embeds [B,S,d] right padded; lens = valid lengths. Returns fp32 logits [B,S,V]."""
    B, S, d = inputs_embeds.shape
    H = cfg["num_attention_heads"]
    hd = d // H
    eps = cfg.get("rms_norm_eps", 1e-5)
    theta = cfg.get("rope_theta", 10000.0)
    lens = [S] * B if lens is None else list(lens)
    pos = torch.arange(S)
    cos, sin = rope_cos_sin(pos, hd, theta)  # [S, hd]
    key_ok = torch.arange(S)[None, :] < torch.tensor(lens)[:, None]  # [B,S]
    causal = torch.tril(torch.ones(S, S, dtype=torch.bool))
    allowed = causal[None, None] & key_ok[:, None, None, :]
    h = inputs_embeds.float()
    for i in range(cfg["num_hidden_layers"]):
        p = f"model.layers.{i}."
        x = rms_norm(h, sd[p + "input_layernorm.weight"], eps)
        q = F.linear(x, sd[p + "self_attn.q_proj.weight"].float()).view(B, S, H, hd).transpose(1, 2)
        k = F.linear(x, sd[p + "self_attn.k_proj.weight"].float()).view(B, S, H, hd).transpose(1, 2)
        v = F.linear(x, sd[p + "self_attn.v_proj.weight"].float()).view(B, S, H, hd).transpose(1, 2)
        q = q * cos + rotate_half(q) * sin
        k = k * cos + rotate_half(k) * sin
        s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        s = s.masked_fill(~allowed, torch.finfo(torch.float32).min)
        a = torch.softmax(s, -1) @ v
        a = a.transpose(1, 2).reshape(B, S, d)
        h = h + F.linear(a, sd[p + "self_attn.o_proj.weight"].float())
        x = rms_norm(h, sd[p + "post_attention_layernorm.weight"], eps)
        g = F.silu(F.linear(x, sd[p + "mlp.gate_proj.weight"].float())) * F.linear(x, sd[p + "mlp.up_proj.weight"].float())
        h = h + F.linear(g, sd[p + "mlp.down_proj.weight"].float())
    h = rms_norm(h, sd["model.norm.weight"], eps)
    return F.linear(h, sd["lm_head.weight"].float())


# ------------------------------------------------------------------ CLIP ViT (LanguageBind)
def _act(name):
    return {"gelu": F.gelu, "quick_gelu": lambda t: t * torch.sigmoid(1.702 * t), "relu": F.relu}[name]


def _mha(x, sd, p, H):
    """HF CLIPAttention: q scaled by hd^-0.5, softmax, out_proj (all with bias)."""
    B, S, d = x.shape
    hd = d // H
    q = F.linear(x, sd[p + "q_proj.weight"].float(), sd[p + "q_proj.bias"].float()) * hd ** -0.5
    k = F.linear(x, sd[p + "k_proj.weight"].float(), sd[p + "k_proj.bias"].float())
    v = F.linear(x, sd[p + "v_proj.weight"].float(), sd[p + "v_proj.bias"].float())
    q, k, v = (t.view(B, S, H, hd).transpose(1, 2) for t in (q, k, v))
    a = torch.softmax(q @ k.transpose(-1, -2), -1) @ v
    a = a.transpose(1, 2).reshape(B, S, d)
    return F.linear(a, sd[p + "out_proj.weight"].float(), sd[p + "out_proj.bias"].float())


def clip_vit_hidden(sd, prefix, cfg, pixels, select_layer=-2):
    """hidden_states[select_layer]: [B,1+np,d] for [B,3,H,W]; [B,T,1+np,d] for [B,3,T,H,W]."""
    d, H, L = cfg["hidden_size"], cfg["num_attention_heads"], cfg["num_hidden_layers"]
    eps = cfg.get("layer_norm_eps", 1e-5)
    act = _act(cfg.get("hidden_act", "gelu"))
    time_attn = cfg.get("add_time_attn", False)
    g = lambda n: sd[prefix + n].float()
    if pixels.dim() == 5:
        B, _, T = pixels.shape[:3]
        px = pixels.permute(0, 2, 1, 3, 4).reshape(B * T, pixels.shape[1], *pixels.shape[3:])
    else:
        B, T, px = pixels.shape[0], 1, pixels
    pe = F.conv2d(px.float(), g("embeddings.patch_embedding.weight"), stride=cfg["patch_size"]).flatten(2).transpose(1, 2)
    cls = g("embeddings.class_embedding").expand(pe.shape[0], 1, -1)
    h = torch.cat([cls, pe], 1) + g("embeddings.position_embedding.weight")[None]
    h = F.layer_norm(h, (d,), g("pre_layrnorm.weight"), g("pre_layrnorm.bias"), eps)
    states = [h]
    n = h.shape[1]
    for i in range(L):
        p = prefix + f"encoder.layers.{i}."
        gl = lambda s: sd[p + s].float()
        if time_attn:
            if T != 1:
                h = (h.view(B, T, n, d) + gl("temporal_embedding")[0, :T][None, :, None, :]).view(B * T, n, d)
            r = h
            x = h.view(B, T, n, d).transpose(1, 2).reshape(B * n, T, d)
            x = F.layer_norm(x, (d,), gl("temporal_layer_norm1.weight"), gl("temporal_layer_norm1.bias"), eps)
            x = _mha(x, sd, p + "temporal_attn.", H)
            h = r + x.view(B, n, T, d).transpose(1, 2).reshape(B * T, n, d)
            if (p + "temporal_mlp.fc1.weight") in sd:
                r = h
                x = F.layer_norm(h, (d,), gl("temporal_layer_norm2.weight"), gl("temporal_layer_norm2.bias"), eps)
                x = F.linear(act(F.linear(x, gl("temporal_mlp.fc1.weight"), gl("temporal_mlp.fc1.bias"))),
                             gl("temporal_mlp.fc2.weight"), gl("temporal_mlp.fc2.bias"))
                h = r + x
        r = h
        x = F.layer_norm(h, (d,), gl("layer_norm1.weight"), gl("layer_norm1.bias"), eps)
        h = r + _mha(x, sd, p + "self_attn.", H)
        r = h
        x = F.layer_norm(h, (d,), gl("layer_norm2.weight"), gl("layer_norm2.bias"), eps)
        h = r + F.linear(act(F.linear(x, gl("mlp.fc1.weight"), gl("mlp.fc1.bias"))), gl("mlp.fc2.weight"), gl("mlp.fc2.bias"))
        states.append(h)
    out = states[select_layer]
    return out.view(B, T, n, d) if pixels.dim() == 5 else out


def projector(sd, prefix, x, depth=2):
    h = x.float()
    for i in range(depth):
        h = F.linear(h, sd[prefix + f"{2 * i}.weight"].float(), sd[prefix + f"{2 * i}.bias"].float())
        if i < depth - 1:
            h = F.gelu(h)
    return h


def region_extractor(sd, prefix, feats, regions, image_size=224):
    """feats [B,S,C]; regions list of [x1,y1,x2,y2] -> [B,1,out]. x indexes rows (layer.py:83)."""
    b, s, c = feats.shape
    g = int(math.sqrt(s))
    masks = []
    for bbox in regions:
        m = torch.zeros((image_size, image_size))
        x1, y1, x2, y2 = bbox
        m[int(x1):int(x2), int(y1):int(y2)] = 1
        masks.append(m)
    m = torch.stack(masks)[:, None]
    f = feats.float().reshape(b, g, g, c).permute(0, 3, 1, 2)
    m = F.interpolate(m, size=(g, g), mode="bilinear", align_corners=False)
    m = (m > 0).float()
    den = m.sum(dim=(-1, -2), keepdim=True) + 1e-8
    pooled = torch.einsum("bchw,bqhw->bqc", f, m / den).reshape(b, c)
    h = pooled
    for i in range(3):
        h = F.linear(h, sd[prefix + f"region_linear.layers.{i}.weight"].float(), sd[prefix + f"region_linear.layers.{i}.bias"].float())
        if i < 2:
            h = F.relu(h)
    loc = torch.tensor([[float(v) for v in r] for r in regions])
    l = F.relu(F.linear(loc, sd[prefix + "loc_encoder.loc_encoder.0.weight"].float(), sd[prefix + "loc_encoder.loc_encoder.0.bias"].float()))
    l = F.linear(l, sd[prefix + "loc_encoder.loc_encoder.2.weight"].float(), sd[prefix + "loc_encoder.loc_encoder.2.bias"].float())
    return (h + l)[:, None]


# ------------------------------------------------------------------ splice + whole model
def encode_all(sd, cfgs, images, regions):
    """Per `images` entry: projected features (list of [n,d]; videos contribute T entries) and region
    features (same length, None where absent) — llava_arch.py:234-290 / 414-446."""
    use_regions = regions is not None and len(regions) > 0
    feats, rfeats = [], []
    for i, im in enumerate(images):
        if im.ndim == 3:
            hid = clip_vit_hidden(sd, "model.image_tower.image_tower.", cfgs["vision"], im[None].float())[:, 1:]
            feats.append(projector(sd, "model.mm_projector.", hid)[0])
            rfeats.append(region_extractor(sd, "model.region_extractor.", hid, [regions[i]], cfgs["vision"]["image_size"])[0]
                          if use_regions else None)
        else:
            hid = clip_vit_hidden(sd, "model.video_tower.video_tower.", cfgs["video"], im[None].float())[:, :, 1:]
            pf = projector(sd, "model.mm_projector.", hid)[0]
            for t in range(pf.shape[0]):
                feats.append(pf[t])
                rfeats.append(None)
    return feats, rfeats


def splice(sd, input_ids, attention_mask, feats, rfeats, use_regions, max_len=None, padding_side="right"):
    """Returns (inputs_embeds [B,S,d] fp32, lens) following llava_arch.py:300-372 / 470-556."""
    E = sd["model.embed_tokens.weight"].float()
    rows = []
    cur = 0
    for b in range(input_ids.shape[0]):
        ids = input_ids[b]
        if attention_mask is not None:
            ids = ids[attention_mask[b].bool()]
        ids = ids.tolist()
        if sum(1 for t in ids if t == IMAGE_TOKEN_INDEX) == 0:
            rows.append(E[torch.tensor(ids, dtype=torch.long)])
            cur += 1
            continue
        parts = []
        for t in ids:
            if t == IMAGE_TOKEN_INDEX:
                parts.append(feats[cur].float())
                cur += 1
            elif t == OBJS_TOKEN_INDEX and use_regions:
                parts.append(rfeats[cur - 1].float())
            else:
                parts.append(E[t][None])
        rows.append(torch.cat(parts, 0))
    if max_len is not None:
        rows = [r[:max_len] for r in rows]
    S = max(r.shape[0] for r in rows)
    out = torch.zeros((len(rows), S, E.shape[1]))
    lens = []
    for b, r in enumerate(rows):
        n = r.shape[0]
        lens.append(n)
        if padding_side == "left":
            out[b, S - n:] = r
        else:
            out[b, :n] = r
    return out, lens


def vitron_logits(sd, cfgs, input_ids, images, regions=None, attention_mask=None):
    """Reference forward(input_ids, images=, regions=).logits for right padding: [B,S,V] fp32."""
    use_regions = regions is not None and len(regions) > 0
    feats, rfeats = encode_all(sd, cfgs, images, regions)
    emb, lens = splice(sd, input_ids, attention_mask, feats, rfeats, use_regions, cfgs.get("max_len"))
    return llama_forward(sd, cfgs["llm"], emb, lens), lens


def greedy_generate(sd, cfgs, input_ids, images, regions=None, max_new_tokens=8, attention_mask=None):
    """Cache-less greedy loop; returns (tokens [B,T] int64, top2_gap [B,T] fp32 logit margins)."""
    use_regions = regions is not None and len(regions) > 0
    feats, rfeats = encode_all(sd, cfgs, images, regions)
    emb, lens = splice(sd, input_ids, attention_mask, feats, rfeats, use_regions, cfgs.get("max_len"))
    E = sd["model.embed_tokens.weight"].float()
    B = emb.shape[0]
    seqs = [emb[b, :lens[b]] for b in range(B)]
    toks, gaps = [], []
    for _ in range(max_new_tokens):
        S = max(s.shape[0] for s in seqs)
        x = torch.zeros((B, S, E.shape[1]))
        for b, s in enumerate(seqs):
            x[b, :s.shape[0]] = s
        ls = [s.shape[0] for s in seqs]
        lg = llama_forward(sd, cfgs["llm"], x, ls)
        last = torch.stack([lg[b, ls[b] - 1] for b in range(B)])
        top2 = last.topk(2, -1).values
        nxt = last.argmax(-1)
        toks.append(nxt)
        gaps.append(top2[:, 0] - top2[:, 1])
        seqs = [torch.cat([s, E[nxt[b]][None]], 0) for b, s in enumerate(seqs)]
    return torch.stack(toks, 1), torch.stack(gaps, 1)


# ------------------------------------------------------------------ cached decoder (CPU baseline timing)
class LlamaCPU:
    """Same arithmetic as llama_forward with the reference's KV cache (tuple of (k, v) grown by
    torch.cat, HF 4.31 LlamaAttention) so that prefill + incremental decode can be timed on the host
    cores as the `cpu_baseline` / `--impl reference` arm of bench.py. Equal to llama_forward
    (tests/test_oracle_cpu.py::test_cached_decoder_equals_full_recompute)."""

    def __init__(self, sd, cfg):
        self.sd, self.cfg = sd, cfg
        self.kv = None
        self.pos = 0

    def _layers(self, h, pos0):
        cfg, sd = self.cfg, self.sd
        B, S, d = h.shape
        H = cfg["num_attention_heads"]
        hd = d // H
        eps, theta = cfg.get("rms_norm_eps", 1e-5), cfg.get("rope_theta", 10000.0)
        cos, sin = rope_cos_sin(torch.arange(pos0, pos0 + S), hd, theta)
        for i in range(cfg["num_hidden_layers"]):
            p = f"model.layers.{i}."
            x = rms_norm(h, sd[p + "input_layernorm.weight"], eps)
            q = F.linear(x, sd[p + "self_attn.q_proj.weight"]).view(B, S, H, hd).transpose(1, 2)
            k = F.linear(x, sd[p + "self_attn.k_proj.weight"]).view(B, S, H, hd).transpose(1, 2)
            v = F.linear(x, sd[p + "self_attn.v_proj.weight"]).view(B, S, H, hd).transpose(1, 2)
            q = q * cos + rotate_half(q) * sin
            k = k * cos + rotate_half(k) * sin
            if self.kv[i] is not None:
                k = torch.cat([self.kv[i][0], k], 2)
                v = torch.cat([self.kv[i][1], v], 2)
            self.kv[i] = (k, v)
            s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
            T = k.shape[2]
            if S > 1:
                allowed = torch.arange(T)[None, :] <= (torch.arange(S)[:, None] + (T - S))
                s = s.masked_fill(~allowed, torch.finfo(torch.float32).min)
            a = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, S, d)
            h = h + F.linear(a, sd[p + "self_attn.o_proj.weight"])
            x = rms_norm(h, sd[p + "post_attention_layernorm.weight"], eps)
            g = F.silu(F.linear(x, sd[p + "mlp.gate_proj.weight"])) * F.linear(x, sd[p + "mlp.up_proj.weight"])
            h = h + F.linear(g, sd[p + "mlp.down_proj.weight"])
        return h

    def prefill(self, inputs_embeds):
        self.kv = [None] * self.cfg["num_hidden_layers"]
        h = self._layers(inputs_embeds.float(), 0)
        self.pos = inputs_embeds.shape[1]
        h = rms_norm(h[:, -1:], self.sd["model.norm.weight"], self.cfg.get("rms_norm_eps", 1e-5))
        return F.linear(h, self.sd["lm_head.weight"])[:, 0]

    def step(self, tokens):
        h = self.sd["model.embed_tokens.weight"][tokens][:, None].float()
        h = self._layers(h, self.pos)
        self.pos += 1
        h = rms_norm(h, self.sd["model.norm.weight"], self.cfg.get("rms_norm_eps", 1e-5))
        return F.linear(h, self.sd["lm_head.weight"])[:, 0]
