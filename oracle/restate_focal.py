"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of SEEM's FocalNet backbone.
Plain torch functional code over a state dict; never imported by the product package.

Follows modules/SEEM/demo_code/xdecoder/backbone/focal.py: PatchEmbed :287-338 (conv embed: stem k7 s4 p2,
downsample k3 s2 p1, input right/bottom zero-padded to a multiple of patch_size), FocalModulation.forward
:91-118, FocalModulationBlock.forward :172-203 (post-LN order when use_postln, layerscale gamma_1/gamma_2),
BasicLayer.forward :275-284, FocalNet.forward :567-590 (per-stage output norm{i}, NCHW outputs res2..res5).

Parity status: PINNED — tests/golden/focal_tiny.pt is produced by the unmodified FocalNet class (imported
through oracle/refshim.setup_seem; timm's DropPath/to_2tuple/trunc_normal_ stubs stated there) and
tests/test_oracle_cpu.py::test_focal_restatement_matches_* compare this file against it and against the live
class on a second configuration.
"""
import torch
import torch.nn.functional as F

FOCAL_L = dict(embed_dim=192, depths=(2, 2, 18, 2), focal_levels=(4, 4, 4, 4), focal_windows=(3, 3, 3, 3),
               mlp_ratio=4.0, patch_size=4, use_conv_embed=True, use_postln=True, use_postln_in_modulation=False,
               scaling_modulator=True, use_layerscale=True, patch_norm=True, out_indices=(0, 1, 2, 3))
"""configs/seem/seem_focall_lang.yaml:29-47 (FocalNet-L as SEEM uses it)."""


def _ln(x, sd, p, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"].float(), sd[p + "bias"].float(), eps)


def patch_embed(sd, p, x, patch_size, use_conv_embed, is_stem, has_norm):
    """PatchEmbed.forward (focal.py:322-338): NCHW in, NCHW out."""
    _, _, H, W = x.shape
    if W % patch_size != 0:
        x = F.pad(x, (0, patch_size - W % patch_size))
    if H % patch_size != 0:
        x = F.pad(x, (0, 0, 0, patch_size - H % patch_size))
    if use_conv_embed:
        stride, pad = (4, 2) if is_stem else (2, 1)
    else:
        stride, pad = patch_size, 0
    x = F.conv2d(x, sd[p + "proj.weight"].float(), sd[p + "proj.bias"].float(), stride=stride, padding=pad)
    if has_norm:
        Wh, Ww = x.shape[2], x.shape[3]
        x = _ln(x.flatten(2).transpose(1, 2), sd, p + "norm.")
        x = x.transpose(1, 2).reshape(-1, x.shape[-1], Wh, Ww)
    return x


def focal_modulation(sd, p, x, focal_level, focal_window, focal_factor=2, use_postln_in_modulation=False,
                     scaling_modulator=False):
    """FocalModulation.forward (focal.py:91-118): x [B, H, W, C] -> [B, H, W, C]."""
    C = x.shape[-1]
    x = F.linear(x, sd[p + "f.weight"].float(), sd[p + "f.bias"].float()).permute(0, 3, 1, 2).contiguous()
    q, ctx, gates = torch.split(x, (C, C, focal_level + 1), 1)
    ctx_all = 0
    for l in range(focal_level):
        k = focal_factor * l + focal_window
        ctx = F.gelu(F.conv2d(ctx, sd[p + f"focal_layers.{l}.0.weight"].float(), None, padding=k // 2, groups=C))
        ctx_all = ctx_all + ctx * gates[:, l:l + 1]
    ctx_global = F.gelu(ctx.mean(2, keepdim=True).mean(3, keepdim=True))
    ctx_all = ctx_all + ctx_global * gates[:, focal_level:]
    if scaling_modulator:
        ctx_all = ctx_all / (focal_level + 1)
    x_out = q * F.conv2d(ctx_all, sd[p + "h.weight"].float(), sd[p + "h.bias"].float())
    x_out = x_out.permute(0, 2, 3, 1).contiguous()
    if use_postln_in_modulation:
        x_out = _ln(x_out, sd, p + "ln.")
    return F.linear(x_out, sd[p + "proj.weight"].float(), sd[p + "proj.bias"].float())


def focal_block(sd, p, x, H, W, cfg, level, window):
    """FocalModulationBlock.forward (focal.py:172-203): x [B, H*W, C]."""
    B, L, C = x.shape
    shortcut = x
    if not cfg["use_postln"]:
        x = _ln(x, sd, p + "norm1.")
    x = focal_modulation(sd, p + "modulation.", x.view(B, H, W, C), level, window,
                         use_postln_in_modulation=cfg["use_postln_in_modulation"],
                         scaling_modulator=cfg["scaling_modulator"]).view(B, H * W, C)
    if cfg["use_postln"]:
        x = _ln(x, sd, p + "norm1.")
    g1 = sd[p + "gamma_1"].float() if cfg["use_layerscale"] else 1.0
    g2 = sd[p + "gamma_2"].float() if cfg["use_layerscale"] else 1.0
    x = shortcut + g1 * x

    def mlp(t):
        t = F.gelu(F.linear(t, sd[p + "mlp.fc1.weight"].float(), sd[p + "mlp.fc1.bias"].float()))
        return F.linear(t, sd[p + "mlp.fc2.weight"].float(), sd[p + "mlp.fc2.bias"].float())

    if cfg["use_postln"]:
        return x + g2 * _ln(mlp(x), sd, p + "norm2.")
    return x + g2 * mlp(_ln(x, sd, p + "norm2."))


def focalnet_forward(sd, x, cfg=FOCAL_L, prefix=""):
    """FocalNet.forward (focal.py:567-590): x [B, 3, H, W] -> {"res2".."res5": [B, C_i, H_i, W_i]}."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    x = patch_embed(sd, "patch_embed.", x.float(), cfg["patch_size"], cfg["use_conv_embed"], True, cfg["patch_norm"])
    Wh, Ww = x.shape[2], x.shape[3]
    x = x.flatten(2).transpose(1, 2)
    outs = {}
    n = len(cfg["depths"])
    for i in range(n):
        H, W = Wh, Ww
        for j in range(cfg["depths"][i]):
            x = focal_block(sd, f"layers.{i}.blocks.{j}.", x, H, W, cfg, cfg["focal_levels"][i], cfg["focal_windows"][i])
        x_out = x
        if i < n - 1:
            xr = x.transpose(1, 2).reshape(x.shape[0], x.shape[-1], H, W)
            xd = patch_embed(sd, f"layers.{i}.downsample.", xr, 2, cfg["use_conv_embed"], False, True)
            Wh, Ww = xd.shape[2], xd.shape[3]
            x = xd.flatten(2).transpose(1, 2)
        if i in cfg["out_indices"]:
            o = _ln(x_out, sd, f"norm{i}.")
            outs[f"res{i + 2}"] = o.view(-1, H, W, o.shape[-1]).permute(0, 3, 1, 2).contiguous()
    return outs
