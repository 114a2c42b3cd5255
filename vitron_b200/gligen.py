"""GLIGEN gated self-attention block on the vitron_b200 kernels.

Drop-in for modules/GLIGEN/demo/gligen/ldm/modules/attention.py: `GatedSelfAttentionDense`
:285-314, `SelfAttention` :192-257, `CrossAttention` :109-186, `FeedForward`/`GEGLU` :45-72 and
`BasicTransformerBlock` :317-349 — same constructor arguments, forward signatures and state-dict
names (`linear`, `attn.to_{q,k,v}`, `attn.to_out.0`, `ff.net.0.proj`, `ff.net.2`, `norm1/2`,
`alpha_attn`, `alpha_dense`).

    x += tanh(alpha_attn) * SelfAttn(LN([x ; Linear(objs)]))[:, :N_visual]
    x += tanh(alpha_dense) * GEGLU-FF(LN(x))

Concat-free: the two LayerNorms write straight into one [B, N+30, C] buffer (no torch.cat), K/V
come from all N+30 rows, queries only from the N visual rows (the reference computes and drops the
30 grounding-token outputs); tanh-gate + residual are the GEMM epilogue (alpha, residual).
Head dims 40 / 80 / 160 run on the padded 48 / 80 / 160 attention kernels.
"""
import math

import torch

from . import ops

BF16 = torch.bfloat16


def _g(sd, name, dev):
    return sd[name].detach().to(device=dev, dtype=BF16).contiguous()


class FeedForward:
    """FeedForward(dim, glu=True): GEGLU(dim, 4 dim) -> Linear(4 dim, dim)."""

    def __init__(self, dim, dim_out=None, mult=4, glu=True, dropout=0., device="cuda"):
        if not glu:
            raise NotImplementedError("GLIGEN uses glu=True everywhere")
        self.dim, self.device = dim, torch.device(device)

    def load_state_dict(self, sd, prefix=""):
        pw, pb = _g(sd, prefix + "net.0.proj.weight", self.device), _g(sd, prefix + "net.0.proj.bias", self.device)
        inner = pw.shape[0] // 2
        self.w1 = ops.pack_glu_weight(pw[:inner], pw[inner:])
        self.b1 = ops.pack_glu_weight(pb[:inner], pb[inner:])
        self.w2, self.b2 = _g(sd, prefix + "net.2.weight", self.device), _g(sd, prefix + "net.2.bias", self.device)
        return self

    def forward(self, x, residual=None, alpha=1.0):
        h = ops.gemm(x, self.w1, bias=self.b1, glu=ops.GLU_GEGLU)
        return ops.gemm(h, self.w2, bias=self.b2, residual=residual, alpha=alpha)

    __call__ = forward


class SelfAttention:
    def __init__(self, query_dim, heads=8, dim_head=64, dropout=0., device="cuda"):
        self.heads, self.dim_head, self.device = heads, dim_head, torch.device(device)

    def load_state_dict(self, sd, prefix=""):
        d = self.device
        self.wqkv = torch.cat([_g(sd, prefix + f"to_{n}.weight", d) for n in "qkv"], 0).contiguous()
        self.wo, self.bo = _g(sd, prefix + "to_out.0.weight", d), _g(sd, prefix + "to_out.0.bias", d)
        return self

    def forward(self, x, n_query=None, residual=None, alpha=1.0):
        """x [B, N, C]; only the first n_query rows are queried (default all)."""
        B, N, C = x.shape
        H, D = self.heads, self.dim_head
        nq = N if n_query is None else n_query
        qkv = ops.gemm(x.reshape(B * N, C), self.wqkv).view(B, N, 3, H, D)
        att = ops.attention(qkv[:, :nq, 0], qkv[:, :, 1], qkv[:, :, 2], scale=D ** -0.5)
        res = None if residual is None else residual.reshape(B * nq, -1)
        return ops.gemm(att.view(B * nq, H * D), self.wo, bias=self.bo, residual=res, alpha=alpha).view(B, nq, -1)

    __call__ = forward


class CrossAttention:
    def __init__(self, query_dim, key_dim, value_dim, heads=8, dim_head=64, dropout=0, device="cuda"):
        self.heads, self.dim_head, self.device = heads, dim_head, torch.device(device)

    def load_state_dict(self, sd, prefix=""):
        d = self.device
        self.wq, self.wk, self.wv = (_g(sd, prefix + f"to_{n}.weight", d) for n in "qkv")
        self.wo, self.bo = _g(sd, prefix + "to_out.0.weight", d), _g(sd, prefix + "to_out.0.bias", d)
        return self

    def forward(self, x, key, value, mask=None, residual=None):
        B, N, C = x.shape
        M = key.shape[1]
        H, D = self.heads, self.dim_head
        q = ops.gemm(x.reshape(B * N, C), self.wq).view(B, N, H, D)
        k = ops.gemm(key.reshape(B * M, -1).to(BF16).contiguous(), self.wk).view(B, M, H, D)
        v = ops.gemm(value.reshape(B * M, -1).to(BF16).contiguous(), self.wv).view(B, M, H, D)
        am = None
        if mask is not None:  # [B, M] bool, True = keep (fill_inf_from_mask :125-131)
            am = (~mask.bool()).view(B, 1, 1, M).expand(B, 1, N, M).contiguous()
        att = ops.attention(q, k, v, scale=D ** -0.5, mask=am)
        res = None if residual is None else residual.reshape(B * N, -1)
        return ops.gemm(att.view(B * N, H * D), self.wo, bias=self.bo, residual=res).view(B, N, -1)

    __call__ = forward


class GatedSelfAttentionDense:
    def __init__(self, query_dim, context_dim, n_heads, d_head, device="cuda"):
        self.query_dim, self.context_dim = query_dim, context_dim
        self.device = torch.device(device)
        self.attn = SelfAttention(query_dim, n_heads, d_head, device=device)
        self.ff = FeedForward(query_dim, glu=True, device=device)
        self.scale = 1
        self.alpha_attn = 0.0
        self.alpha_dense = 0.0

    def load_state_dict(self, sd, prefix=""):
        d = self.device
        self.lw, self.lb = _g(sd, prefix + "linear.weight", d), _g(sd, prefix + "linear.bias", d)
        self.attn.load_state_dict(sd, prefix + "attn.")
        self.ff.load_state_dict(sd, prefix + "ff.")
        self.n1 = (_g(sd, prefix + "norm1.weight", d), _g(sd, prefix + "norm1.bias", d))
        self.n2 = (_g(sd, prefix + "norm2.weight", d), _g(sd, prefix + "norm2.bias", d))
        self.alpha_attn = float(sd[prefix + "alpha_attn"])
        self.alpha_dense = float(sd[prefix + "alpha_dense"])
        return self

    @torch.no_grad()
    def forward(self, x, objs):
        """x [B, N_visual, C], objs [B, n_objs, context_dim] -> [B, N_visual, C]."""
        B, N, C = x.shape
        no = objs.shape[1]
        x = x.to(BF16).contiguous()
        o = ops.gemm(objs.reshape(B * no, -1).to(BF16).contiguous(), self.lw, bias=self.lb).view(B, no, C)
        buf = torch.empty((B, N + no, C), dtype=BF16, device=x.device)
        for b in range(B):  # LayerNorm is row-wise: normalise both sources directly into one buffer
            ops.layernorm(x[b], *self.n1, 1e-5, out=buf[b, :N])
            ops.layernorm(o[b], *self.n1, 1e-5, out=buf[b, N:])
        x = self.attn(buf, n_query=N, residual=x, alpha=self.scale * math.tanh(self.alpha_attn))
        h = ops.layernorm(x.reshape(B * N, C), *self.n2, 1e-5)
        return self.ff(h, residual=x.reshape(B * N, C), alpha=self.scale * math.tanh(self.alpha_dense)).view(B, N, C)

    __call__ = forward


class BasicTransformerBlock:
    """attn1(LN x)+x -> fuser(x, objs) -> attn2(LN x, ctx, ctx)+x -> ff(LN x)+x  (attention.py:344-349)."""

    def __init__(self, query_dim, key_dim, value_dim, n_heads, d_head, fuser_type="gatedSA", use_checkpoint=False,
                 device="cuda"):
        if fuser_type != "gatedSA":
            raise NotImplementedError("only the gatedSA fuser is on the Vitron path")
        self.device = torch.device(device)
        self.attn1 = SelfAttention(query_dim, n_heads, d_head, device=device)
        self.ff = FeedForward(query_dim, glu=True, device=device)
        self.attn2 = CrossAttention(query_dim, key_dim, value_dim, n_heads, d_head, device=device)
        self.fuser = GatedSelfAttentionDense(query_dim, key_dim, n_heads, d_head, device=device)

    def load_state_dict(self, sd, prefix=""):
        d = self.device
        self.attn1.load_state_dict(sd, prefix + "attn1.")
        self.attn2.load_state_dict(sd, prefix + "attn2.")
        self.ff.load_state_dict(sd, prefix + "ff.")
        self.fuser.load_state_dict(sd, prefix + "fuser.")
        self.norms = [(_g(sd, prefix + f"norm{i}.weight", d), _g(sd, prefix + f"norm{i}.bias", d)) for i in (1, 2, 3)]
        return self

    @torch.no_grad()
    def forward(self, x, context, objs):
        B, N, C = x.shape
        x = x.to(BF16).contiguous()
        x = self.attn1(ops.layernorm(x, *self.norms[0], 1e-5), residual=x)
        x = self.fuser(x, objs)
        x = self.attn2(ops.layernorm(x, *self.norms[1], 1e-5), context, context, residual=x)
        h = ops.layernorm(x.reshape(B * N, C), *self.norms[2], 1e-5)
        return self.ff(h, residual=x.reshape(B * N, C)).view(B, N, C)

    __call__ = forward
