"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of GLIGEN's gated self-attention block
(modules/GLIGEN/demo/gligen/ldm/modules/attention.py: SelfAttention.forward_plain :206-225,
CrossAttention.forward_plain :133-155, GEGLU/FeedForward :45-72, GatedSelfAttentionDense :285-314,
BasicTransformerBlock._forward :344-349). PINNED against tests/golden/gligen_tiny.pt (generated
from the unmodified reference file) and the live reference in tests/test_oracle_cpu.py."""
import torch
import torch.nn.functional as F


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"].float(), sd[p + "bias"].float())


def _mha(q_in, k_in, v_in, sd, p, heads, mask=None):
    q = F.linear(q_in, sd[p + "to_q.weight"].float())
    k = F.linear(k_in, sd[p + "to_k.weight"].float())
    v = F.linear(v_in, sd[p + "to_v.weight"].float())
    B, N, HC = q.shape
    C = HC // heads
    q, k, v = (t.view(B, t.shape[1], heads, C).transpose(1, 2) for t in (q, k, v))
    sim = q @ k.transpose(-1, -2) * C ** -0.5
    if mask is not None:
        sim = sim.masked_fill(~mask.bool()[:, None, None, :], -torch.finfo(sim.dtype).max)
    out = (sim.softmax(-1) @ v).transpose(1, 2).reshape(B, N, HC)
    return F.linear(out, sd[p + "to_out.0.weight"].float(), sd[p + "to_out.0.bias"].float())


def _ff(x, sd, p):
    h = F.linear(x, sd[p + "net.0.proj.weight"].float(), sd[p + "net.0.proj.bias"].float())
    a, gate = h.chunk(2, dim=-1)
    return F.linear(a * F.gelu(gate), sd[p + "net.2.weight"].float(), sd[p + "net.2.bias"].float())


def gated_self_attention_dense(sd, prefix, x, objs, heads, scale=1.0):
    p = prefix if (prefix == "" or prefix.endswith(".")) else prefix + "."
    n = x.shape[1]
    o = F.linear(objs.float(), sd[p + "linear.weight"].float(), sd[p + "linear.bias"].float())
    x = x.float()
    h = _ln(torch.cat([x, o], dim=1), sd, p + "norm1.")
    x = x + scale * torch.tanh(sd[p + "alpha_attn"].float()) * _mha(h, h, h, sd, p + "attn.", heads)[:, :n]
    return x + scale * torch.tanh(sd[p + "alpha_dense"].float()) * _ff(_ln(x, sd, p + "norm2."), sd, p + "ff.")


def basic_transformer_block(sd, prefix, x, context, objs, heads, scale=1.0):
    p = prefix if (prefix == "" or prefix.endswith(".")) else prefix + "."
    x = x.float()
    h = _ln(x, sd, p + "norm1.")
    x = _mha(h, h, h, sd, p + "attn1.", heads) + x
    x = gated_self_attention_dense(sd, p + "fuser.", x, objs, heads, scale)   # `scale` = evaluator.py set_alpha_scale
    c = context.float()
    x = _mha(_ln(x, sd, p + "norm2."), c, c, sd, p + "attn2.", heads) + x
    return _ff(_ln(x, sd, p + "norm3."), sd, p + "ff.") + x
