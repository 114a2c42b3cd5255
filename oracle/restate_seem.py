"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of SEEM's pixel decoder and mask decoder for
task='seg', extra={} (panoptic path). Plain torch; never imported by the product package.

Follows modules/SEEM/demo_code/xdecoder: body/encoder/transformer_encoder_fpn.py:23-308
(detectron2 `Conv2d(norm=GN, activation=relu)` == conv(bias=False) -> GroupNorm(32) -> ReLU, NORM "GN",
configs/seem/seem_focall_lang.yaml:55), body/transformer_blocks.py:154-232 (post-norm encoder layer),
body/decoder/seem.py:29-189,395-586, body/decoder/utils/utils.py:18-32,
body/decoder/utils/attention_data_struct.py:173-187,250-264, body/decoder/utils/attn.py:296-316
(bool mask -> -inf, softmax(...).nan_to_num()), modules/position_encoding.py:18-52,
language/vlpencoder.py:293-299.

Parity status: PINNED at module level. The unmodified reference classes TransformerEncoderPixelDecoder and
MultiScaleMaskedTransformerDecoder are imported through oracle/refshim.setup_seem (which states the few
detectron2 / fvcore / timm layer definitions they use, absent from this image) and this restatement
reproduces their outputs to 3e-4: tests/golden/seem_tiny.pt (oracle/gen_golden.py::gen_seem) and a live
comparison on a second configuration (tests/test_oracle_cpu.py::test_seem_restatement_matches_*), in
addition to the piecewise pins of multi_head_attention_forward, PositionEmbeddingSine, prepare_features
and AttentionDataStruct's mask rule.
"""
import math

import torch
import torch.nn.functional as F


def position_embedding_sine(x, num_pos_feats, temperature=10000, scale=2 * math.pi):
    """PositionEmbeddingSine(normalize=True).forward(x, mask=None) -> [B, 2*npf, H, W]."""
    not_mask = torch.ones((x.size(0), x.size(2), x.size(3)), dtype=torch.bool)
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps = 1e-6
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def mha(query, key, value, sd, p, heads, attn_mask=None):
    """[L, B, C] layout like nn.MultiheadAttention; bool attn_mask [B*heads, L, S], True = masked."""
    w, b = sd[p + "in_proj_weight"].float(), sd[p + "in_proj_bias"].float()
    C = w.shape[1]
    q = F.linear(query, w[:C], b[:C])
    k = F.linear(key, w[C:2 * C], b[C:2 * C])
    v = F.linear(value, w[2 * C:], b[2 * C:])
    L, B, _ = q.shape
    S = k.shape[0]
    hd = C // heads
    q = q.contiguous().view(L, B * heads, hd).transpose(0, 1) * hd ** -0.5
    k = k.contiguous().view(S, B * heads, hd).transpose(0, 1)
    v = v.contiguous().view(S, B * heads, hd).transpose(0, 1)
    s = torch.bmm(q, k.transpose(1, 2))
    if attn_mask is not None:
        s = s.masked_fill(attn_mask, float("-inf"))
    a = F.softmax(s, dim=-1).nan_to_num()
    o = torch.bmm(a, v).transpose(0, 1).contiguous().view(L, B, C)
    return F.linear(o, sd[p + "out_proj.weight"].float(), sd[p + "out_proj.bias"].float())


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"].float(), sd[p + "bias"].float())


def _conv_gn_relu(x, sd, p, relu=True, pad=1):
    y = F.conv2d(x, sd[p + "weight"].float(), None, padding=pad)
    y = F.group_norm(y, 32, sd[p + "norm.weight"].float(), sd[p + "norm.bias"].float(), 1e-5)
    return F.relu(y) if relu else y


def pixel_decoder_forward(sd, features, prefix="", nheads=8, enc_layers=6, in_features=("res2", "res3", "res4", "res5")):
    """TransformerEncoderPixelDecoder.forward_features -> (mask_features, enc_features, multi_scale[3])."""
    g = lambda n: sd[prefix + n].float()
    multi, y, enc = [], None, None
    L = len(in_features)
    for idx, f in enumerate(in_features[::-1]):
        x = features[f].float()
        k = L - idx
        if idx == 0:
            t = F.conv2d(x, g("input_proj.weight"), g("input_proj.bias"))
            bs, c, h, w = t.shape
            pos = position_embedding_sine(x, c // 2).flatten(2).permute(2, 0, 1)
            src = t.flatten(2).permute(2, 0, 1)
            for i in range(enc_layers):
                p = prefix + f"transformer.encoder.layers.{i}."
                qk = src + pos
                src = _ln(src + mha(qk, qk, src, sd, p + "self_attn.", nheads), sd, p + "norm1.")
                ff = F.linear(F.relu(F.linear(src, sd[p + "linear1.weight"].float(), sd[p + "linear1.bias"].float())),
                              sd[p + "linear2.weight"].float(), sd[p + "linear2.bias"].float())
                src = _ln(src + ff, sd, p + "norm2.")
            enc = src.permute(1, 2, 0).view(bs, c, h, w)
            y = _conv_gn_relu(enc, sd, prefix + f"layer_{k}.")
        else:
            cur = _conv_gn_relu(x, sd, prefix + f"adapter_{k}.", relu=False, pad=0)
            y = cur + F.interpolate(y, size=cur.shape[-2:], mode="nearest")
            y = _conv_gn_relu(y, sd, prefix + f"layer_{k}.")
        if len(multi) < 3:
            multi.append(y)
    mf = F.conv2d(y, g("mask_features.weight"), g("mask_features.bias"), padding=1)
    return mf, enc, multi


def prediction_heads(sd, prefix, output, mask_features, target_size, heads, t_emb, logit_scale):
    """forward_prediction_heads (seem.py:555-586); output [Q, B, C]."""
    dec = _ln(output, sd, prefix + "decoder_norm.").transpose(0, 1)
    class_embed = dec @ sd[prefix + "class_embed"].float()
    outputs_class = None
    if t_emb is not None:
        v = class_embed / (class_embed.norm(dim=-1, keepdim=True) + 1e-7)
        outputs_class = math.exp(logit_scale) * v @ t_emb.float().unsqueeze(0).transpose(1, 2)
    me = dec
    for i in range(3):
        me = F.linear(me, sd[prefix + f"mask_embed.layers.{i}.weight"].float(), sd[prefix + f"mask_embed.layers.{i}.bias"].float())
        if i < 2:
            me = F.relu(me)
    outputs_mask = torch.einsum("bqc,bchw->bqhw", me, mask_features.float())
    am = F.interpolate(outputs_mask, size=target_size, mode="bilinear", align_corners=False)
    am = (am.sigmoid().flatten(2).unsqueeze(1).repeat(1, heads, 1, 1).flatten(0, 1) < 0.5).bool()
    return dict(attn_mask=am, predictions_class=outputs_class, predictions_mask=outputs_mask, predictions_maskemb=me)


def mask_decoder_forward(sd, x, mask_features, prefix="", heads=8, num_layers=9, t_emb=None, logit_scale=0.0):
    """MultiScaleMaskedTransformerDecoder.forward(task='seg', extra={})."""
    g = lambda n: sd[prefix + n].float()
    C = g("query_feat.weight").shape[1]
    src, pos, size_list = [], [], []
    for i in range(3):
        size_list.append(tuple(x[i].shape[-2:]))
        pos.append(position_embedding_sine(x[i], C // 2).flatten(2).permute(2, 0, 1))
        src.append((x[i].float().flatten(2) + g("level_embed.weight")[i][None, :, None]).permute(2, 0, 1))
    bs = src[0].shape[1]
    query_embed = g("query_embed.weight").unsqueeze(1).repeat(1, bs, 1)
    output = g("query_feat.weight").unsqueeze(1).repeat(1, bs, 1)
    results = [prediction_heads(sd, prefix, output, mask_features, size_list[0], heads, t_emb, logit_scale)]
    for i in range(num_layers):
        lvl = i % 3
        am = results[-1]["attn_mask"].clone()
        am[torch.where(am.sum(-1) == am.shape[-1])] = False  # AttentionDataStruct.cross_attn_mask :187
        p = prefix + f"transformer_cross_attention_layers.{i}."
        t2 = mha(output + query_embed, src[lvl] + pos[lvl], src[lvl], sd, p + "multihead_attn.", heads, am)
        output = _ln(output + t2, sd, p + "norm.")
        p = prefix + f"transformer_self_attention_layers.{i}."
        qk = output + query_embed
        output = _ln(output + mha(qk, qk, output, sd, p + "self_attn.", heads, None), sd, p + "norm.")
        p = prefix + f"transformer_ffn_layers.{i}."
        ff = F.linear(F.relu(F.linear(output, sd[p + "linear1.weight"].float(), sd[p + "linear1.bias"].float())),
                      sd[p + "linear2.weight"].float(), sd[p + "linear2.bias"].float())
        output = _ln(output + ff, sd, p + "norm.")
        results.append(prediction_heads(sd, prefix, output, mask_features, size_list[(i + 1) % 3], heads, t_emb, logit_scale))
    names = {"predictions_class": "pred_logits", "predictions_mask": "pred_masks", "predictions_maskemb": "pred_maskembs"}
    out = {v: results[-1][k] for k, v in names.items()}
    out["aux_outputs"] = [{v: r[k] for k, v in names.items()} for r in results[:-1]]
    out["attn_masks"] = [r["attn_mask"] for r in results]
    return out


def seem_shapes(in_channels=(192, 384, 768, 1536), C=512, ffn=2048, Q=101, enc_layers=6, dec_layers=9, dim_proj=512):
    """Reference parameter names / shapes of sem_seg_head.{pixel_decoder,predictor} on the seg path."""
    s = {}
    p = "pixel_decoder."
    s[p + "input_proj.weight"], s[p + "input_proj.bias"] = [C, in_channels[-1], 1, 1], [C]

    def mha_(q):
        s[q + "in_proj_weight"], s[q + "in_proj_bias"] = [3 * C, C], [3 * C]
        s[q + "out_proj.weight"], s[q + "out_proj.bias"] = [C, C], [C]

    def ffn_(q):
        s[q + "linear1.weight"], s[q + "linear1.bias"] = [ffn, C], [ffn]
        s[q + "linear2.weight"], s[q + "linear2.bias"] = [C, ffn], [C]

    def norm_(q):
        s[q + "weight"], s[q + "bias"] = [C], [C]

    for i in range(enc_layers):
        q = p + f"transformer.encoder.layers.{i}."
        mha_(q + "self_attn.")
        ffn_(q)
        norm_(q + "norm1.")
        norm_(q + "norm2.")
    for idx, cin in enumerate(in_channels):
        k = idx + 1
        s[p + f"layer_{k}.weight"] = [C, C, 3, 3]
        norm_(p + f"layer_{k}.norm.")
        if idx != len(in_channels) - 1:
            s[p + f"adapter_{k}.weight"] = [C, cin, 1, 1]
            norm_(p + f"adapter_{k}.norm.")
    s[p + "mask_features.weight"], s[p + "mask_features.bias"] = [C, C, 3, 3], [C]
    p = "predictor."
    for i in range(dec_layers):
        mha_(p + f"transformer_cross_attention_layers.{i}.multihead_attn.")
        norm_(p + f"transformer_cross_attention_layers.{i}.norm.")
        mha_(p + f"transformer_self_attention_layers.{i}.self_attn.")
        norm_(p + f"transformer_self_attention_layers.{i}.norm.")
        ffn_(p + f"transformer_ffn_layers.{i}.")
        norm_(p + f"transformer_ffn_layers.{i}.norm.")
    norm_(p + "decoder_norm.")
    s[p + "query_feat.weight"], s[p + "query_embed.weight"], s[p + "level_embed.weight"] = [Q, C], [Q, C], [3, C]
    for i in range(3):
        s[p + f"mask_embed.layers.{i}.weight"], s[p + f"mask_embed.layers.{i}.bias"] = [C, C], [C]
    s[p + "class_embed"] = [C, dim_proj]
    return s
