"""OpenCLIP text / image embedders of i2vgen-xl on the vitron_b200 kernels.

Drop-ins for modules/i2vgen-xl/tools/modules/clip_embedder.py: `FrozenOpenCLIPEmbedder` (:13-75, text),
`FrozenOpenCLIPVisualEmbedder` (:78-142, image) and `FrozenOpenCLIPTtxtVisualEmbedder` (:146-228, both — the one
the entrance calls: `y_visual, y_text, y_words = clip_encoder(image=..., text=...)`, inference_i2vgen_entrance.py:166-168).
What is the reference's own code there — token + positional embedding, running all / all-but-one resblocks
(`layer="penultimate"` -> `layer_idx = 1`), `ln_final`, EOS-token pooling `x[arange, text.argmax(-1)] @ text_projection`,
`encode_image` — is mirrored; the transformer arithmetic itself belongs to the third-party `open_clip` package
(ViT-H-14: text width 1024 / 24 layers / 16 heads / 77 tokens / vocab 49408; vision width 1280 / 32 layers /
16 heads / patch 14 / mlp 5120 / output 1024), whose published ResidualAttentionBlock is
x += out_proj(MHA(ln_1(x))), x += c_proj(GELU(c_fc(ln_2(x)))) with a causal mask in the text tower.
State-dict names are open_clip's (`model.transformer.resblocks.N.attn.in_proj_weight`, `model.visual.conv1.weight`, ...).
Tokenisation (`open_clip.tokenize`, a host-side BPE) is not part of this module: `text` is the int token tensor
[B, 77] (a str raises).

Kernels: fused in_proj GEMM (+bias), flash attention (causal for text; head_dim 64 text / 80 vision), out_proj / c_proj
GEMMs with the residual in the epilogue, c_fc GEMM with exact-erf GELU, LayerNorm; the image tower reuses the
patchify + GEMM + fused cls/pos/pre-LN path of the LanguageBind tower.
"""
import torch

from . import ops
from .vision_tower import VisionConfig, VisionTransformerB200

BF16 = torch.bfloat16

VIT_H_14 = dict(embed_dim=1024, text=dict(width=1024, layers=24, heads=16, context_length=77, vocab_size=49408),
                vision=dict(width=1280, layers=32, heads=16, patch_size=14, image_size=224, mlp=5120))
"""open_clip model_configs/ViT-H-14.json."""


def _blocks(sd, prefix, n, dev):
    g = lambda name: sd[prefix + name].detach().to(device=dev, dtype=BF16).contiguous()
    out = []
    for i in range(n):
        p = f"resblocks.{i}."
        out.append(dict(wqkv=g(p + "attn.in_proj_weight"), bqkv=g(p + "attn.in_proj_bias"), wo=g(p + "attn.out_proj.weight"),
                        bo=g(p + "attn.out_proj.bias"), ln1w=g(p + "ln_1.weight"), ln1b=g(p + "ln_1.bias"),
                        ln2w=g(p + "ln_2.weight"), ln2b=g(p + "ln_2.bias"), w1=g(p + "mlp.c_fc.weight"), b1=g(p + "mlp.c_fc.bias"),
                        w2=g(p + "mlp.c_proj.weight"), b2=g(p + "mlp.c_proj.bias")))
    return out


def _run_blocks(h, layers, nb, n, heads, causal):
    """h [nb*n, d] bf16, updated in place: pre-LN residual attention blocks (open_clip ResidualAttentionBlock)."""
    d = h.shape[1]
    hd = d // heads
    for L in layers:
        x = ops.layernorm(h, L["ln1w"], L["ln1b"], 1e-5)
        qkv = ops.gemm(x, L["wqkv"], bias=L["bqkv"]).view(nb, n, 3, heads, hd)
        att = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale=hd ** -0.5, causal=causal)
        ops.gemm(att.view(nb * n, d), L["wo"], bias=L["bo"], residual=h, out=h)
        x = ops.layernorm(h, L["ln2w"], L["ln2b"], 1e-5)
        f = ops.gemm(x, L["w1"], bias=L["b1"], act=ops.ACT_GELU)
        ops.gemm(f, L["w2"], bias=L["b2"], residual=h, out=h)
    return h


class _Base:
    LAYERS = ["last", "penultimate"]

    def __init__(self, pretrained=None, arch="ViT-H-14", device="cuda", max_length=77, freeze=True, layer="last",
                 arch_cfg=None, **kwargs):
        if layer not in self.LAYERS:
            raise AssertionError(layer)
        self.device = torch.device(device)
        self.max_length = max_length
        self.layer = layer
        self.layer_idx = 0 if layer == "last" else 1
        self.cfg = arch_cfg or VIT_H_14
        self.text = None
        self.visual = None

    def freeze(self):
        return self

    # ---- text tower -------------------------------------------------------------------------------
    def _load_text(self, sd, prefix="model."):
        dev, t = self.device, self.cfg["text"]
        g = lambda name: sd[prefix + name].detach().to(device=dev, dtype=BF16).contiguous()
        self.text = dict(tok=g("token_embedding.weight"), pos=g("positional_embedding"),
                         blocks=_blocks(sd, prefix + "transformer.", t["layers"], dev),
                         lnw=g("ln_final.weight"), lnb=g("ln_final.bias"),
                         proj_t=g("text_projection").t().contiguous())   # [embed_dim, width]: x @ text_projection

    @torch.no_grad()
    def encode_with_transformer(self, text):
        """clip_embedder.py:49-56 / :190-198: tokens [B, 77] -> (EOS-pooled projection [B, embed], ln_final(x) [B, 77, width])."""
        if isinstance(text, (str, list)) and not torch.is_tensor(text):
            raise ValueError("pass open_clip token ids [B, 77]; tokenisation is host-side and not part of this module")
        if self.text is None:
            raise RuntimeError("load_state_dict() first")
        t = self.text
        tokens = text.to(self.device).long()
        B, S = tokens.shape
        heads = self.cfg["text"]["heads"]
        x = torch.empty((B * S, t["tok"].shape[1]), dtype=BF16, device=self.device)
        srcmap = tokens.reshape(-1).to(torch.int32)
        ops.splice_multimodal(t["tok"], None, srcmap, out=x)                       # embedding gather
        x = ops.add(x, t["pos"][:S].contiguous())                                  # + positional_embedding (broadcast over B)
        n_run = len(t["blocks"]) - self.layer_idx                                  # "penultimate": skip the last resblock
        _run_blocks(x, t["blocks"][:n_run], B, S, heads, causal=True)
        x = ops.layernorm(x, t["lnw"], t["lnb"], 1e-5).view(B, S, -1)
        eos = x[torch.arange(B, device=self.device), tokens.argmax(dim=-1)].contiguous()
        xt = ops.gemm(eos, t["proj_t"])
        return xt, x

    # ---- image tower ------------------------------------------------------------------------------
    def _load_visual(self, sd, prefix="model.visual."):
        dev, v = self.device, self.cfg["vision"]
        cfg = VisionConfig(hidden_size=v["width"], intermediate_size=v["mlp"], num_hidden_layers=v["layers"],
                           num_attention_heads=v["heads"], image_size=v["image_size"], patch_size=v["patch_size"], hidden_act="gelu")
        vit = VisionTransformerB200(cfg, dev)
        hf = {"embeddings.patch_embedding.weight": sd[prefix + "conv1.weight"], "embeddings.class_embedding": sd[prefix + "class_embedding"],
              "embeddings.position_embedding.weight": sd[prefix + "positional_embedding"],
              "pre_layrnorm.weight": sd[prefix + "ln_pre.weight"], "pre_layrnorm.bias": sd[prefix + "ln_pre.bias"]}
        d = v["width"]
        for i in range(v["layers"]):
            p, q = prefix + f"transformer.resblocks.{i}.", f"encoder.layers.{i}."
            w, b = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
            for j, n in enumerate("qkv"):
                hf[q + f"self_attn.{n}_proj.weight"], hf[q + f"self_attn.{n}_proj.bias"] = w[j * d:(j + 1) * d], b[j * d:(j + 1) * d]
            hf[q + "self_attn.out_proj.weight"], hf[q + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"]
            hf[q + "layer_norm1.weight"], hf[q + "layer_norm1.bias"] = sd[p + "ln_1.weight"], sd[p + "ln_1.bias"]
            hf[q + "layer_norm2.weight"], hf[q + "layer_norm2.bias"] = sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]
            hf[q + "mlp.fc1.weight"], hf[q + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]
            hf[q + "mlp.fc2.weight"], hf[q + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"]
        vit.load_state_dict(hf)
        g = lambda name: sd[prefix + name].detach().to(device=dev, dtype=BF16).contiguous()
        self.visual = dict(vit=vit, lnw=g("ln_post.weight"), lnb=g("ln_post.bias"), proj_t=g("proj").t().contiguous())

    @torch.no_grad()
    def encode_image(self, image):
        """open_clip CLIP.encode_image = visual(image): CLS token -> ln_post -> @ proj  -> [B, embed_dim]."""
        if self.visual is None:
            raise RuntimeError("load_state_dict() first (this embedder was built without the image tower)")
        v = self.visual
        h = v["vit"].forward_hidden(image.to(self.device), select_layer=v["vit"].config.num_hidden_layers)
        cls = ops.layernorm(h[:, 0].contiguous(), v["lnw"], v["lnb"], 1e-5)
        return ops.gemm(cls, v["proj_t"])


class FrozenOpenCLIPEmbedder(_Base):
    """Text only (cfg.embedder default, tools/modules/config.py:129-133): forward(tokens) -> ln_final(x) [B, 77, 1024]."""

    def load_state_dict(self, sd, strict=True):
        self._load_text(sd)
        return self

    def forward(self, text):
        return self.encode_with_transformer(text)[1]

    def encode(self, text):
        return self(text)

    __call__ = forward


class FrozenOpenCLIPVisualEmbedder(_Base):
    def load_state_dict(self, sd, strict=True):
        self._load_visual(sd)
        return self

    def forward(self, image):
        return self.encode_image(image)

    __call__ = forward


class FrozenOpenCLIPTtxtVisualEmbedder(_Base):
    """(sic) forward(image=None, text=tokens) -> (xi [B, 1024] | None, xt [B, 1024], x [B, 77, 1024])."""

    def load_state_dict(self, sd, strict=True):
        self._load_text(sd)
        self._load_visual(sd)
        return self

    def forward(self, image=None, text=None):
        xi = self.encode_image(image) if image is not None else None
        xt, x = self.encode_with_transformer(text)
        return xi, xt, x

    def encode(self, text):
        return self(text=text)

    __call__ = forward
