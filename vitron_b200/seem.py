"""SEEM pixel decoder + mask decoder on the vitron_b200 kernels (panoptic / 'seg' path).

Drop-ins for modules/SEEM/demo_code/xdecoder:
  * `TransformerEncoderPixelDecoder.forward_features(features)` — body/encoder/
    transformer_encoder_fpn.py:194-303 (FPN: 1x1 lateral + nearest-upsample add + 3x3 conv+GN+ReLU;
    6 post-norm encoder layers on res5 — body/transformer_blocks.py:154-232; 3x3 mask_features conv)
  * `MultiScaleMaskedTransformerDecoder.forward(x, mask_features, mask=None, target_queries=None,
    target_vlp=None, task='seg', extra={})` — body/decoder/seem.py:395-586 with
    CrossAttentionLayer/SelfAttentionLayer/FFNLayer :29-189 (post-norm), `prepare_features`
    (utils/utils.py:18-32), PositionEmbeddingSine (modules/position_encoding.py:18-52), the
    bool-mask rule of AttentionDataStruct.cross_attn_mask (attention_data_struct.py:173-187) and
    LanguageEncoder.compute_similarity (language/vlpencoder.py:293-299)
  * `XDecoderHead.layers` = pixel_decoder + predictor (body/xdecoder_head.py:103-118)
State-dict names are the reference's. Interactive prompts (spatial / grounding / visual / audio
`extra` keys) are not on the BASELINE path and raise NotImplementedError (DESIGN.md "next").

B200 design: NHWC bf16 throughout; the K/V projections of the memory for the 3 layers that share a
feature level are ONE grouped GEMM per level (N = 3 x 512); the mask head `einsum('bqc,bchw->bqhw')`
is a GEMM against the NHWC mask_features viewed as [HW, C]; the next layer's bool attention mask
(bilinear resize, sigmoid < 0.5, fully-masked-row reset) is one fused kernel with the threshold
decision in fp32; masked cross-attention never materialises Q x HW scores.
"""
import math

import torch
import torch.nn.functional as F

from . import ops

BF16 = torch.bfloat16


def position_embedding_sine(h, w, num_pos_feats, device, temperature=10000, scale=2 * math.pi):
    """PositionEmbeddingSine(normalize=True) for an unmasked h x w map -> [h*w, 2*num_pos_feats] fp32."""
    y_embed = torch.arange(1, h + 1, dtype=torch.float32, device=device)[:, None].expand(h, w)
    x_embed = torch.arange(1, w + 1, dtype=torch.float32, device=device)[None, :].expand(h, w)
    eps = 1e-6
    y_embed = y_embed / (y_embed[-1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=device)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    pos_x = x_embed[:, :, None] / dim_t
    pos_y = y_embed[:, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, 0::2].sin(), pos_x[:, :, 1::2].cos()), dim=3).flatten(2)
    pos_y = torch.stack((pos_y[:, :, 0::2].sin(), pos_y[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((pos_y, pos_x), dim=2).reshape(h * w, 2 * num_pos_feats)


def _nhwc(x):
    """NCHW (any float dtype) -> contiguous NHWC bf16."""
    return x.permute(0, 2, 3, 1).to(BF16).contiguous()


def _nearest_to(y, h, w):
    n, hy, wy, c = y.shape
    if h == 2 * hy and w == 2 * wy:
        return ops.upsample2x_nhwc(y)
    iy = torch.div(torch.arange(h, device=y.device) * hy, h, rounding_mode="floor")
    ix = torch.div(torch.arange(w, device=y.device) * wy, w, rounding_mode="floor")
    return y[:, iy][:, :, ix].contiguous()


class _MHA:
    """nn.MultiheadAttention weights split for fused use: q/k share their input in every SEEM layer."""

    def __init__(self, sd, p, dev):
        g = lambda n: sd[p + n].detach().to(device=dev, dtype=BF16).contiguous()
        w, b = g("in_proj_weight"), g("in_proj_bias")
        d = w.shape[1]
        self.d = d
        self.wq, self.wk, self.wv = w[:d].contiguous(), w[d:2 * d].contiguous(), w[2 * d:].contiguous()
        self.bq, self.bk, self.bv = b[:d].contiguous(), b[d:2 * d].contiguous(), b[2 * d:].contiguous()
        self.wqk, self.bqk = w[:2 * d].contiguous(), b[:2 * d].contiguous()
        self.wo, self.bo = g("out_proj.weight"), g("out_proj.bias")


class TransformerEncoderPixelDecoder:
    def __init__(self, in_channels=(192, 384, 768, 1536), conv_dim=512, mask_dim=512, nheads=8, dim_feedforward=2048,
                 enc_layers=6, in_features=("res2", "res3", "res4", "res5"), device="cuda"):
        self.in_features = list(in_features)
        self.in_channels = list(in_channels)
        self.conv_dim, self.mask_dim, self.nheads, self.enc_layers = conv_dim, mask_dim, nheads, enc_layers
        self.device = torch.device(device)
        self.maskformer_num_feature_levels = 3
        self._pos = {}

    def load_state_dict(self, sd, prefix=""):
        dev = self.device
        g = lambda n: sd[prefix + n].detach().to(device=dev, dtype=BF16).contiguous()
        L = len(self.in_features)
        w = {"input_proj": (g("input_proj.weight").reshape(self.conv_dim, -1).contiguous(), g("input_proj.bias"))}
        w["enc"] = []
        for i in range(self.enc_layers):
            p = f"transformer.encoder.layers.{i}."
            w["enc"].append(dict(att=_MHA(sd, prefix + p + "self_attn.", dev), l1=(g(p + "linear1.weight"), g(p + "linear1.bias")),
                                 l2=(g(p + "linear2.weight"), g(p + "linear2.bias")), n1=(g(p + "norm1.weight"), g(p + "norm1.bias")),
                                 n2=(g(p + "norm2.weight"), g(p + "norm2.bias"))))
        for idx in range(L):
            k = idx + 1
            w[f"layer_{k}"] = (ops.pack_conv_weight(g(f"layer_{k}.weight")), g(f"layer_{k}.norm.weight"), g(f"layer_{k}.norm.bias"))
            if idx != L - 1:
                aw = g(f"adapter_{k}.weight")
                w[f"adapter_{k}"] = (aw.reshape(aw.shape[0], -1).contiguous(), g(f"adapter_{k}.norm.weight"), g(f"adapter_{k}.norm.bias"))
        w["mask_features"] = (ops.pack_conv_weight(g("mask_features.weight")), g("mask_features.bias"))
        self.w = w
        return self

    def _pe(self, h, w):
        key = (h, w)
        if key not in self._pos:
            self._pos[key] = position_embedding_sine(h, w, self.conv_dim // 2, self.device).to(BF16).contiguous()
        return self._pos[key]

    def _out_conv(self, y, name):
        cw, gw, gb = self.w[name]
        return ops.groupnorm_nhwc(ops.conv_nhwc(y, cw, 3, 3), gw, gb, 32, 1e-5, act=ops.ACT_RELU)

    @torch.no_grad()
    def forward_features(self, features):
        """features: {'res2'..'res5': [B, C, H, W]} -> (mask_features, transformer_encoder_features,
        multi_scale_features[3]) as NCHW-shaped views of NHWC bf16 storage."""
        C, H = self.conv_dim, self.nheads
        hd = C // H
        multi = []
        y = None
        enc_feat = None
        for idx, f in enumerate(self.in_features[::-1]):
            x = _nhwc(features[f].to(self.device))
            n, h, w, cin = x.shape
            k = len(self.in_features) - idx
            if idx == 0:
                src = ops.gemm(x.view(n * h * w, cin), self.w["input_proj"][0], bias=self.w["input_proj"][1])
                pos = self._pe(h, w)
                for lw in self.w["enc"]:
                    a = lw["att"]
                    qk_in = ops.add(src, pos)
                    qk = ops.gemm(qk_in, a.wqk, bias=a.bqk).view(n, h * w, 2, H, hd)
                    v = ops.gemm(src, a.wv, bias=a.bv).view(n, h * w, H, hd)
                    att = ops.attention(qk[:, :, 0], qk[:, :, 1], v, scale=hd ** -0.5)
                    ops.gemm(att.view(n * h * w, C), a.wo, bias=a.bo, residual=src, out=src)
                    ops.layernorm(src, *lw["n1"], 1e-5, out=src)
                    ff = ops.gemm(src, lw["l1"][0], bias=lw["l1"][1], act=ops.ACT_RELU)
                    ops.gemm(ff, lw["l2"][0], bias=lw["l2"][1], residual=src, out=src)
                    ops.layernorm(src, *lw["n2"], 1e-5, out=src)
                enc_feat = src.view(n, h, w, C)
                y = self._out_conv(enc_feat, f"layer_{k}")
            else:
                aw, gw, gb = self.w[f"adapter_{k}"]
                cur = ops.gemm(x.view(n * h * w, cin), aw).view(n, h, w, C)
                cur = ops.groupnorm_nhwc(cur, gw, gb, 32, 1e-5)
                y = ops.add(cur, _nearest_to(y, h, w))
                y = self._out_conv(y, f"layer_{k}")
            if len(multi) < self.maskformer_num_feature_levels:
                multi.append(y)
        mf = ops.conv_nhwc(y, self.w["mask_features"][0], 3, 3, bias=self.w["mask_features"][1])
        nchw = lambda t: t.permute(0, 3, 1, 2)
        return nchw(mf), nchw(enc_feat), [nchw(m) for m in multi]


class MultiScaleMaskedTransformerDecoder:
    def __init__(self, hidden_dim=512, dim_proj=512, num_queries=101, nheads=8, dim_feedforward=2048, dec_layers=9,
                 mask_dim=512, device="cuda"):
        self.hidden_dim, self.dim_proj, self.num_queries = hidden_dim, dim_proj, num_queries
        self.num_heads, self.num_layers, self.mask_dim = nheads, dec_layers, mask_dim
        self.num_feature_levels = 3
        self.device = torch.device(device)
        self.text_embeddings = None
        self.logit_scale = 0.0
        self._pos = {}

    def set_text_embeddings(self, t_emb, logit_scale):
        """Stand-in for lang_encoder.default_text_embeddings / logit_scale (vlpencoder.py:293-299)."""
        self.text_embeddings = t_emb.detach().to(device=self.device, dtype=BF16).contiguous()
        self.logit_scale = float(logit_scale)

    def load_state_dict(self, sd, prefix=""):
        dev = self.device
        g = lambda n: sd[prefix + n].detach().to(device=dev, dtype=BF16).contiguous()
        L = self.num_layers
        self.cross = [dict(att=_MHA(sd, prefix + f"transformer_cross_attention_layers.{i}.multihead_attn.", dev),
                           n=(g(f"transformer_cross_attention_layers.{i}.norm.weight"), g(f"transformer_cross_attention_layers.{i}.norm.bias")))
                      for i in range(L)]
        self.selfa = [dict(att=_MHA(sd, prefix + f"transformer_self_attention_layers.{i}.self_attn.", dev),
                           n=(g(f"transformer_self_attention_layers.{i}.norm.weight"), g(f"transformer_self_attention_layers.{i}.norm.bias")))
                      for i in range(L)]
        self.ffn = [dict(l1=(g(f"transformer_ffn_layers.{i}.linear1.weight"), g(f"transformer_ffn_layers.{i}.linear1.bias")),
                         l2=(g(f"transformer_ffn_layers.{i}.linear2.weight"), g(f"transformer_ffn_layers.{i}.linear2.bias")),
                         n=(g(f"transformer_ffn_layers.{i}.norm.weight"), g(f"transformer_ffn_layers.{i}.norm.bias"))) for i in range(L)]
        self.decoder_norm = (g("decoder_norm.weight"), g("decoder_norm.bias"))
        self.query_feat, self.query_embed, self.level_embed = g("query_feat.weight"), g("query_embed.weight"), g("level_embed.weight")
        self.mask_embed = [(g(f"mask_embed.layers.{i}.weight"), g(f"mask_embed.layers.{i}.bias")) for i in range(3)]
        self.class_embed_t = g("class_embed").t().contiguous()  # [dim_proj, hidden] as a GEMM weight
        # grouped K / V projection weights per feature level (level l feeds layers l, l+3, l+6)
        self.kgrp, self.vgrp = [], []
        for lvl in range(self.num_feature_levels):
            ids = [i for i in range(L) if i % self.num_feature_levels == lvl]
            self.kgrp.append((ids, torch.cat([self.cross[i]["att"].wk for i in ids], 0).contiguous(),
                              torch.cat([self.cross[i]["att"].bk for i in ids], 0).contiguous()))
            self.vgrp.append((ids, torch.cat([self.cross[i]["att"].wv for i in ids], 0).contiguous(),
                              torch.cat([self.cross[i]["att"].bv for i in ids], 0).contiguous()))
        return self

    def _pe(self, h, w):
        if (h, w) not in self._pos:
            self._pos[(h, w)] = position_embedding_sine(h, w, self.hidden_dim // 2, self.device).to(BF16).contiguous()
        return self._pos[(h, w)]

    def _heads(self, output, mask_rows, mask_hw, target_size, bs):
        """forward_prediction_heads (seem.py:555-586). output [bs*Q, C] rows (batch-major)."""
        Q, C = self.num_queries, self.hidden_dim
        dec = ops.layernorm(output, *self.decoder_norm, 1e-5)
        class_embed = ops.gemm(dec, self.class_embed_t)  # decoder_output @ class_embed
        outputs_class = None
        if self.text_embeddings is not None:
            ce = class_embed.float()
            v = (ce / (ce.norm(dim=-1, keepdim=True) + 1e-7)).to(BF16).contiguous()
            outputs_class = ops.gemm(v, self.text_embeddings, alpha=math.exp(self.logit_scale), out_fp32=True)
            outputs_class = outputs_class.view(bs, Q, -1)
        me = dec
        for i, (w, b) in enumerate(self.mask_embed):
            me = ops.gemm(me, w, bias=b, act=ops.ACT_RELU if i < 2 else ops.ACT_NONE)
        H, W = mask_hw
        outputs_mask = torch.empty((bs, Q, H * W), dtype=torch.float32, device=output.device)
        for b in range(bs):  # einsum('bqc,bchw->bqhw') == mask_embed[b] @ mask_features[b]^T over NHWC rows
            ops.gemm(me[b * Q:(b + 1) * Q], mask_rows[b], out=outputs_mask[b], out_fp32=True)
        outputs_mask = outputs_mask.view(bs, Q, H, W)
        attn_mask = ops.seem_attn_mask(outputs_mask.view(bs * Q, H, W), int(target_size[0]), int(target_size[1]))
        attn_mask = attn_mask.view(bs, 1, Q, -1)  # shared by the heads; rows that were fully masked are already reset
        return dict(attn_mask=attn_mask, predictions_class=outputs_class, predictions_mask=outputs_mask,
                    predictions_caption=class_embed.view(bs, Q, -1), predictions_maskemb=me.view(bs, Q, -1))

    @torch.no_grad()
    def forward(self, x, mask_features, mask=None, target_queries=None, target_vlp=None, task="seg", extra={}):
        if extra:
            raise NotImplementedError("interactive prompts (spatial / grounding / visual / audio) are not on the BASELINE path")
        assert len(x) == self.num_feature_levels
        Q, C, H = self.num_queries, self.hidden_dim, self.num_heads
        hd = C // H
        bs = x[0].shape[0]
        dev = self.device
        # prepare_features: src = feat + level_embed (V input), src + pos (K input); rows are (b, hw)
        size_list, kproj, vproj = [], {}, {}
        for lvl in range(self.num_feature_levels):
            f = _nhwc(x[lvl].to(dev))
            n, h, w, c = f.shape
            size_list.append((h, w))
            src = ops.add(f.view(n * h * w, c), self.level_embed[lvl].contiguous())
            kin = ops.add(src, self._pe(h, w))
            ids, wk, bk = self.kgrp[lvl]
            kall = ops.gemm(kin, wk, bias=bk).view(n, h * w, len(ids), H, hd)
            _, wv, bv = self.vgrp[lvl]
            vall = ops.gemm(src, wv, bias=bv).view(n, h * w, len(ids), H, hd)
            for j, i in enumerate(ids):
                kproj[i], vproj[i] = kall[:, :, j], vall[:, :, j]
        mf = _nhwc(mask_features.to(dev))
        mask_hw = (mf.shape[1], mf.shape[2])
        mask_rows = [mf[b].view(-1, mf.shape[-1]) for b in range(bs)]

        output = self.query_feat.unsqueeze(0).repeat(bs, 1, 1).view(bs * Q, C).contiguous()
        qpos = self.query_embed.unsqueeze(0).repeat(bs, 1, 1).view(bs * Q, C).contiguous()
        results = [self._heads(output, mask_rows, mask_hw, size_list[0], bs)]
        for i in range(self.num_layers):
            lvl = i % self.num_feature_levels
            # masked cross attention (post-norm)
            a = self.cross[i]["att"]
            q = ops.gemm(ops.add(output, qpos), a.wq, bias=a.bq).view(bs, Q, H, hd)
            att = ops.attention(q, kproj[i], vproj[i], scale=hd ** -0.5, mask=results[-1]["attn_mask"])
            ops.gemm(att.view(bs * Q, C), a.wo, bias=a.bo, residual=output, out=output)
            ops.layernorm(output, *self.cross[i]["n"], 1e-5, out=output)
            # self attention among the object queries (mask all-False for task 'seg')
            a = self.selfa[i]["att"]
            qk = ops.gemm(ops.add(output, qpos), a.wqk, bias=a.bqk).view(bs, Q, 2, H, hd)
            v = ops.gemm(output, a.wv, bias=a.bv).view(bs, Q, H, hd)
            att = ops.attention(qk[:, :, 0], qk[:, :, 1], v, scale=hd ** -0.5)
            ops.gemm(att.view(bs * Q, C), a.wo, bias=a.bo, residual=output, out=output)
            ops.layernorm(output, *self.selfa[i]["n"], 1e-5, out=output)
            # FFN
            f1 = ops.gemm(output, self.ffn[i]["l1"][0], bias=self.ffn[i]["l1"][1], act=ops.ACT_RELU)
            ops.gemm(f1, self.ffn[i]["l2"][0], bias=self.ffn[i]["l2"][1], residual=output, out=output)
            ops.layernorm(output, *self.ffn[i]["n"], 1e-5, out=output)
            results.append(self._heads(output, mask_rows, mask_hw, size_list[(i + 1) % self.num_feature_levels], bs))
        # organize_output (attention_data_struct.py:250-264) for the object queries
        names = {"predictions_class": "pred_logits", "predictions_mask": "pred_masks", "predictions_maskemb": "pred_maskembs"}
        out = {v: results[-1][k] for k, v in names.items()}
        out["aux_outputs"] = [{v: r[k] for k, v in names.items()} for r in results[:-1]]
        out["attn_masks"] = [r["attn_mask"] for r in results]  # extra (not in the reference dict): for parity tests
        return out

    __call__ = forward


class XDecoderHead:
    """sem_seg_head: layers(features) = predictor(*pixel_decoder.forward_features(features))
    (body/xdecoder_head.py:100-118)."""

    def __init__(self, pixel_decoder, predictor):
        self.pixel_decoder, self.predictor = pixel_decoder, predictor
        self._graphed = None

    def enable_graph(self, on=True):
        """Replay the whole head (pixel decoder + 9 decoder layers + 10 prediction heads, ~400 launches) from one CUDA graph
        per feature-shape signature. The returned tensors are then static buffers, overwritten by the next call."""
        self._graphed = ops.GraphedCall(lambda **f: self.layers(f)) if on else None
        return self

    def load_state_dict(self, sd, prefix=""):
        self.pixel_decoder.load_state_dict(sd, prefix + "pixel_decoder.")
        self.predictor.load_state_dict(sd, prefix + "predictor.")
        return self

    def forward(self, features, mask=None, target_queries=None, target_vlp=None, task="seg", extra={}):
        if (self._graphed is not None and not extra and mask is None and target_queries is None and target_vlp is None
                and task == "seg" and all(torch.is_tensor(v) and v.is_cuda for v in features.values())):
            return self._graphed(**features)
        return self.layers(features, mask, target_queries, target_vlp, task, extra)

    def layers(self, features, mask=None, target_queries=None, target_vlp=None, task="seg", extra={}):
        mask_features, enc_feats, multi = self.pixel_decoder.forward_features(features)
        return self.predictor(multi, mask_features, mask, target_queries, target_vlp, task, extra)

    __call__ = forward
