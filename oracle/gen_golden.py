"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden [name ...]
Each fixture stores configs, parameter shapes + seed (weights are re-derived with
oracle.weights.seeded_state_dict), inputs and the reference's outputs.
"""
import os
import sys

import torch

from . import refshim
from .weights import seeded_state_dict, shapes_of

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

LLM_TINY = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, vocab_size=320)
VIT_TINY = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=56,
                patch_size=14)


def _load_seeded(model, seed, gain=0.8):
    shapes = shapes_of(model)
    sd = seeded_state_dict(shapes, seed, gain)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    # only non-persistent buffers (rotary inv_freq, position_ids) may be absent
    assert all(("inv_freq" in m or "position_ids" in m) for m in missing), missing
    return shapes


def gen_vitron_llm(seed=11):
    """Image + region + video batch through the reference LlavaLlamaForCausalLM: logits + greedy ids."""
    torch.manual_seed(0)
    model = refshim.build_reference_vitron(LLM_TINY, VIT_TINY, with_video=True, num_frames=4)
    shapes = _load_seeded(model, seed)
    g = torch.Generator().manual_seed(seed)
    V = LLM_TINY["vocab_size"]
    # sample 0: one image with a region; sample 1: one image, no objs token; both right-aligned lengths differ
    img0 = torch.randn((3, 56, 56), generator=g)
    img1 = torch.randn((3, 56, 56), generator=g)
    ids0 = [1] + torch.randint(3, V, (5,), generator=g).tolist() + [-200] + torch.randint(3, V, (4,), generator=g).tolist() + [-300] + torch.randint(3, V, (3,), generator=g).tolist()
    ids1 = [1] + torch.randint(3, V, (3,), generator=g).tolist() + [-200] + torch.randint(3, V, (6,), generator=g).tolist()
    L = max(len(ids0), len(ids1))
    ids = torch.zeros((2, L), dtype=torch.long)
    am = torch.zeros((2, L), dtype=torch.long)
    ids[0, :len(ids0)] = torch.tensor(ids0); am[0, :len(ids0)] = 1
    ids[1, :len(ids1)] = torch.tensor(ids1); am[1, :len(ids1)] = 1
    regions = [[8.0, 12.0, 40.0, 50.0], [0.0, 0.0, 56.0, 56.0]]
    out = {"llm": LLM_TINY, "vit": VIT_TINY, "num_frames": 4, "seed": seed, "shapes": shapes}
    with torch.no_grad():
        r = model(input_ids=ids, attention_mask=am, images=[img0, img1], regions=regions, use_cache=False)
        out["img"] = dict(input_ids=ids, attention_mask=am, images=[img0, img1], regions=regions, logits=r.logits.float())
        # single-sample greedy generation (image + region), cached path of the reference
        gen = model.generate(ids[:1, :len(ids0)], images=[img0], regions=regions[:1], do_sample=False,
                             max_new_tokens=12, use_cache=True)
        out["gen_img"] = dict(input_ids=ids[:1, :len(ids0)], images=[img0], regions=regions[:1], tokens=gen[:, len(ids0):])
        # video sample: 4 frames -> 4 <image> sentinels
        vid = torch.randn((3, 4, 56, 56), generator=g)
        vids = [1] + [-200] * 4 + torch.randint(3, V, (6,), generator=g).tolist()
        vids = torch.tensor([vids])
        rv = model(input_ids=vids, images=[vid], use_cache=False)
        genv = model.generate(vids, images=[vid], do_sample=False, max_new_tokens=8, use_cache=True)
        out["vid"] = dict(input_ids=vids, images=[vid], logits=rv.logits.float(), tokens=genv[:, vids.shape[1]:])
        # tower / adapter level outputs
        it = model.get_model().image_tower
        feats = it(torch.stack([img0, img1]))
        out["tower"] = dict(image_feats=feats.float(),
                            video_feats=model.get_model().video_tower(vid[None]).float(),
                            proj=model.get_model().mm_projector(feats).float(),
                            region=model.get_model().region_extractor(feats, regions).float())
    torch.save(out, os.path.join(OUT, "vitron_llm_tiny.pt"))
    print("vitron_llm_tiny.pt", {k: (tuple(v["logits"].shape) if isinstance(v, dict) and "logits" in v else "") for k, v in out.items() if isinstance(v, dict)})


UNET_TINY = dict(in_dim=4, concat_dim=4, dim=64, y_dim=1024, context_dim=1024, out_dim=4, dim_mult=[1, 2], num_heads=2, head_dim=64,
                 num_res_blocks=1, attn_scales=[1.0, 0.5], num_tokens=4)


def gen_unet(seed=21):
    """UNetSD_I2VGen forward (b=2, f=4, 8x16 latent) + 3 DDIM steps with CFG through the reference."""
    U = refshim.i2vgen_unet_class()
    model = U(**UNET_TINY, dropout=0.1, temporal_attention=True, temporal_attn_times=1, use_checkpoint=False,
              use_fps_condition=True, use_sim_mask=False, training=False, inpainting=True).eval()
    shapes = _load_seeded(model, seed, gain=0.4)
    g = torch.Generator().manual_seed(seed)
    b, f, h, w = 2, 4, 8, 16
    rn = lambda *s: torch.randn(s, generator=g)
    inp = dict(x=rn(b, 4, f, h, w), t=torch.tensor([500, 37]), y=rn(b, 77, 1024), image=rn(b, 1, 1024),
               local_image=rn(b, 4, f, h, w), fps=torch.tensor([8, 16]))
    out = {"cfg": UNET_TINY, "seed": seed, "gain": 0.4, "shapes": shapes, "inputs": inp}
    with torch.no_grad():
        out["out"] = model(**inp).float()
        D = refshim.i2vgen_ddim_class()
        diff = D(schedule="cosine", schedule_param=dict(num_timesteps=1000, cosine_s=0.008, zero_terminal_snr=True),
                 mean_type="v", loss_type="mse", var_type="fixed_small", rescale_timesteps=False, noise_strength=0.1)
        noise = rn(1, 4, f, h, w)
        cond = dict(y=inp["y"][:1], image=inp["image"][:1], local_image=inp["local_image"][:1], fps=inp["fps"][:1])
        unc = dict(y=inp["y"][1:], image=torch.zeros_like(inp["image"][:1]), local_image=inp["local_image"][:1], fps=inp["fps"][:1])
        vid = diff.ddim_sample_loop(noise=noise, model=model, model_kwargs=[cond, unc], guide_scale=9.0, ddim_timesteps=4, eta=0.0)
        out["ddim"] = dict(noise=noise, cond=cond, uncond=unc, guide_scale=9.0, ddim_timesteps=4, out=vid.float())
    torch.save(out, os.path.join(OUT, "unet_tiny.pt"))
    print("unet_tiny.pt", tuple(out["out"].shape), float(out["out"].abs().max()), tuple(vid.shape), float(vid.abs().max()))


def gen_gligen(seed=31):
    """GatedSelfAttentionDense + BasicTransformerBlock(fuser=gatedSA) at the three GLIGEN head sizes."""
    att = refshim.gligen_attention()
    g = torch.Generator().manual_seed(seed)
    ctx = torch.randn((2, 20, 768), generator=g)
    objs = torch.randn((2, 30, 768), generator=g)
    out = {"seed": seed, "context": ctx, "objs": objs, "cases": []}
    for (N, C, heads) in ((64, 320, 8), (36, 640, 8), (16, 1280, 8)):
        dh = C // heads
        blk = att.BasicTransformerBlock(C, 768, 768, heads, dh, "gatedSA", use_checkpoint=False).eval()
        shapes = shapes_of(blk)
        sd = seeded_state_dict(shapes, seed)
        blk.load_state_dict(sd)
        x = torch.randn((2, N, C), generator=g)
        with torch.no_grad():
            fo = blk.fuser(x, objs)
            bo = blk(x, ctx, objs)
        out["cases"].append(dict(N=N, C=C, heads=heads, shapes=shapes, x=x, fuser_out=fo, block_out=bo))
    torch.save(out, os.path.join(OUT, "gligen_tiny.pt"))
    print("gligen_tiny.pt", [tuple(c["block_out"].shape) for c in out["cases"]])


SEEM_TINY = dict(in_channels=(32, 48, 64, 96), C=128, ffn=256, Q=16, enc_layers=2, dec_layers=3, dim_proj=64, heads=2,
                 n_text=8, logit_scale=1.3)


def _seem_attn_arch():
    """ATTENTION_ARCH of configs/seem/seem_focall_lang.yaml:114-139."""
    return {
        "VARIABLE": {"queries": ["object"], "tokens": ["grounding", "spatial", "visual", "audio"]},
        "SELF_ATTENTION": {
            "queries": {"object": ["queries_object", "tokens_grounding", "tokens_spatial", "tokens_visual", "tokens_audio"]},
            "tokens": {"grounding": ["queries_object", "tokens_grounding"], "spatial": ["tokens_spatial"],
                       "visual": ["tokens_visual"], "audio": ["queries_object", "tokens_audio"]}},
        "CROSS_ATTENTION": {"queries": {"object": True},
                            "tokens": {"grounding": False, "spatial": False, "visual": False, "audio": False}},
        "MASKING": ["tokens_spatial", "tokens_grounding", "tokens_visual", "tokens_audio"],
        "DUPLICATION": {"queries": {"grounding": "queries_object", "spatial": "queries_object"}},
        "SPATIAL_MEMORIES": 32,
    }


def build_reference_seem(t=SEEM_TINY, seed=41, prompts=False):
    """The UNMODIFIED TransformerEncoderPixelDecoder + MultiScaleMaskedTransformerDecoder (through the
    detectron2-layer stubs of refshim.setup_seem), seeded weights; returns (pixel_decoder, predictor, sd)."""
    import torch.nn as nn
    PD, MD, ShapeSpec = refshim.seem_classes()
    shape = {f"res{i + 2}": ShapeSpec(channels=c, stride=4 << i) for i, c in enumerate(t["in_channels"])}
    pd = PD(input_shape=shape, transformer_dropout=0.0, transformer_nheads=t["heads"], transformer_dim_feedforward=t["ffn"],
            transformer_enc_layers=t["enc_layers"], transformer_pre_norm=False, conv_dim=t["C"], mask_dim=t["C"],
            mask_on=True, norm="GN").eval()

    class Lang(nn.Module):
        """Only `compute_similarity` of language/vlpencoder.py:293-299 is on the seg path; the text
        embeddings are an input (the text tower is out of scope)."""

        def __init__(self):
            super().__init__()
            self.logit_scale = nn.Parameter(torch.tensor(float(t["logit_scale"])))
            self.register_buffer("default_text_embeddings", torch.zeros(t["n_text"], t["dim_proj"]))

        def compute_similarity(self, v_emb, name="default", fake=False):
            v_emb = v_emb / (v_emb.norm(dim=-1, keepdim=True) + 1e-7)
            t_emb = getattr(self, "{}_text_embeddings".format(name))
            return self.logit_scale.exp() * v_emb @ t_emb.unsqueeze(0).transpose(1, 2)

    # task switch of configs/seem/seem_focall_lang.yaml:60-86 (MASK, SPATIAL enabled; the rest disabled)
    task_switch = {"bbox": False, "mask": True, "spatial": True, "grounding": prompts, "openimage": {"grounding": False, "mask": False},
                   "visual": prompts, "audio": prompts}
    md = MD(Lang(), t["C"], True, hidden_dim=t["C"], dim_proj=t["dim_proj"], num_queries=t["Q"], contxt_len=77,
            nheads=t["heads"], dim_feedforward=t["ffn"], dec_layers=t["dec_layers"], pre_norm=False, mask_dim=t["C"],
            task_switch=task_switch, enforce_input_project=False, max_spatial_len=[512, 512, 512, 512],
            attn_arch=_seem_attn_arch()).eval()
    shapes = {}
    shapes.update({"pixel_decoder." + k: v for k, v in shapes_of(pd).items()})
    shapes.update({"predictor." + k: v for k, v in shapes_of(md).items() if not k.startswith("lang_encoder.")})
    sd = seeded_state_dict(shapes, seed)
    pd.load_state_dict({k[len("pixel_decoder."):]: v for k, v in sd.items() if k.startswith("pixel_decoder.")})
    miss, unexp = md.load_state_dict({k[len("predictor."):]: v for k, v in sd.items() if k.startswith("predictor.")}, strict=False)
    assert not unexp and all(m.startswith("lang_encoder.") for m in miss), (miss, unexp)
    return pd, md, sd, shapes


def gen_seem(seed=41):
    """SEEM pixel decoder + mask decoder, task='seg', extra={} from the unmodified reference classes."""
    t = SEEM_TINY
    pd, md, sd, shapes = build_reference_seem(t, seed)
    g = torch.Generator().manual_seed(seed)
    feats = {f"res{i + 2}": torch.randn((1, c, 32 >> i, 40 >> i), generator=g) for i, c in enumerate(t["in_channels"])}
    t_emb = torch.randn((t["n_text"], t["dim_proj"]), generator=g)
    t_emb = t_emb / t_emb.norm(dim=-1, keepdim=True)
    md.lang_encoder.default_text_embeddings.copy_(t_emb)
    with torch.no_grad():
        mf, enc, multi = pd.forward_features(feats)
        out = md(multi, mf, task="seg", extra={})
    keep = {k: out[k] for k in ("pred_logits", "pred_masks", "pred_maskembs")}
    keep["aux_outputs"] = [{k: a[k] for k in ("pred_logits", "pred_masks", "pred_maskembs")} for a in out["aux_outputs"]]
    fx = dict(seed=seed, cfg=dict(t), shapes=shapes, features=feats, t_emb=t_emb, mask_features=mf, enc_features=enc,
              multi_scale=multi, out=keep, out_keys=sorted(out.keys()))
    torch.save(fx, os.path.join(OUT, "seem_tiny.pt"))
    print("seem_tiny.pt", tuple(mf.shape), tuple(out["pred_masks"].shape), float(out["pred_masks"].abs().max()),
          len(out["aux_outputs"]), sorted(out.keys()))


def gen_seem_prompts(seed=43):
    """SEEM mask decoder with INTERACTIVE prompts from the unmodified reference class (seem.py:398-500 +
    attention_data_struct.py): grounding tokens (text), spatial positive / negative point masks, audio tokens, and the
    `refimg` -> visual-prompt route. Point counts stay below max_spatial_len, so rand_sample draws nothing."""
    t = SEEM_TINY
    pd, md, sd, shapes = build_reference_seem(t, seed, prompts=True)
    g = torch.Generator().manual_seed(seed)
    feats = {f"res{i + 2}": torch.randn((1, c, 32 >> i, 40 >> i), generator=g) for i, c in enumerate(t["in_channels"])}
    t_emb = torch.randn((t["n_text"], t["dim_proj"]), generator=g)
    t_emb = t_emb / t_emb.norm(dim=-1, keepdim=True)
    md.lang_encoder.default_text_embeddings.copy_(t_emb)
    C = t["C"]
    gtok = torch.randn((5, 1, C), generator=g) * 0.5
    atok = torch.randn((3, 1, C), generator=g) * 0.5
    pos_mask = torch.zeros((1, 32, 40), dtype=torch.bool)
    pos_mask[0, 5:9, 10:14] = True
    neg_mask = torch.zeros((1, 32, 40), dtype=torch.bool)
    neg_mask[0, 20:22, 30:33] = True
    cases = {
        "grounding": dict(grounding_tokens=gtok, grounding_nonzero_mask=torch.zeros((1, 5), dtype=torch.bool)),
        "spatial": dict(spatial_query_pos_mask=[pos_mask], spatial_query_neg_mask=[neg_mask]),
        "grounding+spatial+audio": dict(grounding_tokens=gtok, grounding_nonzero_mask=torch.zeros((1, 5), dtype=torch.bool),
                                        spatial_query_pos_mask=[pos_mask], spatial_query_neg_mask=[neg_mask],
                                        audio_tokens=atok, audio_nonzero_mask=torch.zeros((1, 3), dtype=torch.bool)),
    }
    keys = ("pred_logits", "pred_masks", "pred_maskembs", "pred_captions", "pred_pspatials", "pred_nspatials", "pred_pvisuals", "pred_nvisuals")
    outs = {}
    def switch(**on):   # tasks/interactive.py:53-57: the demo resets the task switches per request, then enables what it uses
        for k in ("spatial", "visual", "grounding", "audio"):
            md.task_switch[k] = bool(on.get(k, False))
    with torch.no_grad():
        mf, enc, multi = pd.forward_features(feats)
        for name, extra in cases.items():
            switch(grounding="grounding_tokens" in extra, spatial="spatial_query_pos_mask" in extra, audio="audio_tokens" in extra)
            out = md(multi, mf, task="seg", extra=dict(extra))
            outs[name] = {k: out[k] for k in keys if k in out and out[k] is not None}
            outs[name]["aux0"] = {k: out["aux_outputs"][0][k] for k in keys if k in out["aux_outputs"][0] and out["aux_outputs"][0][k] is not None}
        switch(spatial=True, visual=True)
        ref = md(multi, mf, task="refimg", extra=dict(cases["spatial"]))                # evaluate_referring_image route
        switch(visual=True)
        vis_extra = dict(visual_query_pos=ref["visual_query_pos"], visual_query_neg=ref["visual_query_neg"],
                         src_visual_queries=ref["src_visual_queries"], src_visual_maskings=ref["src_visual_maskings"])
        out = md(multi, mf, task="seg", extra=dict(vis_extra))
        outs["visual"] = {k: out[k] for k in keys if k in out and out[k] is not None}
    fx = dict(seed=seed, cfg=dict(t), shapes=shapes, t_emb=t_emb, mask_features=mf, multi_scale=multi, cases=cases, refimg=ref, out=outs)
    torch.save(fx, os.path.join(OUT, "seem_prompts_tiny.pt"))
    print("seem_prompts_tiny.pt", {k: sorted(v.keys()) for k, v in outs.items()})


FOCAL_TINY = dict(embed_dim=64, depths=(1, 1, 2, 1), focal_levels=(4, 4, 4, 4), focal_windows=(3, 3, 3, 3), mlp_ratio=4.0,
                  patch_size=4, use_conv_embed=True, use_postln=True, use_postln_in_modulation=False,
                  scaling_modulator=True, use_layerscale=True, patch_norm=True, out_indices=(0, 1, 2, 3))


def build_reference_focalnet(cfg, seed):
    """The UNMODIFIED FocalNet (backbone/focal.py:340-590) in eval mode with seeded weights (layerscale gammas are
    seeded at O(0.5) instead of the 1e-4 init, otherwise the blocks would not contribute)."""
    FocalNet = refshim.seem_focalnet_class()
    net = FocalNet(pretrain_img_size=224, patch_size=cfg["patch_size"], in_chans=3, embed_dim=cfg["embed_dim"],
                   depths=list(cfg["depths"]), mlp_ratio=cfg["mlp_ratio"], drop_rate=0.0, drop_path_rate=0.3,
                   patch_norm=cfg["patch_norm"], out_indices=list(cfg["out_indices"]), focal_levels=list(cfg["focal_levels"]),
                   focal_windows=list(cfg["focal_windows"]), use_conv_embed=cfg["use_conv_embed"], use_postln=cfg["use_postln"],
                   use_postln_in_modulation=cfg["use_postln_in_modulation"], scaling_modulator=cfg["scaling_modulator"],
                   use_layerscale=cfg["use_layerscale"])
    net.eval()  # FocalNet.train() returns None (focal.py:592-595), so no chaining
    shapes = shapes_of(net)
    sd = seeded_state_dict(shapes, seed)
    net.load_state_dict(sd)
    return net, sd, shapes


def gen_focal(seed=51):
    """FocalNet backbone (SEEM's FocalNet-L block structure at tiny widths) from the unmodified reference class."""
    net, sd, shapes = build_reference_focalnet(FOCAL_TINY, seed)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((1, 3, 64, 96), generator=g)
    with torch.no_grad():
        outs = net(x)
    torch.save(dict(seed=seed, cfg=dict(FOCAL_TINY), shapes=shapes, x=x, outs=outs), os.path.join(OUT, "focal_tiny.pt"))
    print("focal_tiny.pt", {k: (tuple(v.shape), round(float(v.abs().max()), 3)) for k, v in outs.items()})


VAE_TINY = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=64, ch_mult=(1, 2, 2),
                num_res_blocks=1, attn_resolutions=[], dropout=0.0)


def build_reference_vae(ddconfig, seed):
    AE = refshim.i2vgen_autoencoder_class()
    ae = AE(ddconfig=dict(ddconfig), embed_dim=4).eval()
    shapes = _load_seeded(ae, seed, gain=0.8)
    return ae, seeded_state_dict(shapes, seed, 0.8), shapes


def gen_vae(seed=61):
    """AutoencoderKL encode (posterior moments) + decode through the unmodified reference class."""
    ae, sd, shapes = build_reference_vae(VAE_TINY, seed)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((2, 3, 32, 48), generator=g)
    z = torch.randn((2, 4, 8, 12), generator=g)
    with torch.no_grad():
        post = ae.encode(x)
        dec = ae.decode(z)
    torch.save(dict(seed=seed, ddconfig=dict(VAE_TINY), shapes=shapes, x=x, z=z, mean=post.mean, logvar=post.logvar,
                    std=post.std, dec=dec), os.path.join(OUT, "vae_tiny.pt"))
    print("vae_tiny.pt", tuple(post.mean.shape), float(post.mean.abs().max()), tuple(dec.shape), float(dec.abs().max()))


GLIGEN_UNET_TINY = dict(image_size=16, in_channels=4, model_channels=128, out_channels=4, num_res_blocks=1,
                        attention_resolutions=[2, 1], channel_mult=[1, 2, 2], num_heads=2, transformer_depth=1, context_dim=96,
                        positive_len=96, fuser_type="gatedSA", use_checkpoint=False)


def build_reference_gligen_unet(cfg, seed):
    import contextlib
    import io
    U = refshim.gligen_unet_class()
    with contextlib.redirect_stdout(io.StringIO()):
        net = U(**cfg).eval()
    shapes = _load_seeded(net, seed, gain=0.6)
    return net, seeded_state_dict(shapes, seed, 0.6), shapes


def gligen_unet_inputs(cfg, g, b=2, n_obj=5, hw=(16, 16), inpaint=False):
    rn = lambda *s: torch.randn(s, generator=g)
    inp = dict(x=rn(b, cfg["in_channels"], *hw), timesteps=torch.tensor([801, 37][:b]), context=rn(b, 9, cfg["context_dim"]),
               boxes=torch.rand((b, n_obj, 4), generator=g), masks=(torch.rand((b, n_obj), generator=g) > 0.4).float(),
               text_embeddings=rn(b, n_obj, cfg["positive_len"]))
    if inpaint:
        inp["inpainting_extra_input"] = rn(b, cfg["in_channels"] + 1, *hw)
    return inp


def gen_gligen_unet(seed=71):
    """GLIGEN's grounded UNetModel forward (reference prints suppressed) through the unmodified class."""
    import contextlib
    import io
    net, sd, shapes = build_reference_gligen_unet(GLIGEN_UNET_TINY, seed)
    inp = gligen_unet_inputs(GLIGEN_UNET_TINY, torch.Generator().manual_seed(seed))
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        out = net(dict(inp))
    torch.save(dict(seed=seed, gain=0.6, cfg=dict(GLIGEN_UNET_TINY), shapes=shapes, inputs=inp, out=out),
               os.path.join(OUT, "gligen_unet_tiny.pt"))
    print("gligen_unet_tiny.pt", tuple(out.shape), float(out.abs().max()), len(shapes), "tensors")


GENERATORS = {"seem_prompts": gen_seem_prompts, "vitron_llm": gen_vitron_llm, "unet": gen_unet, "gligen": gen_gligen, "seem": gen_seem, "focal": gen_focal, "vae": gen_vae, "gligen_unet": gen_gligen_unet}


def main(argv):
    os.makedirs(OUT, exist_ok=True)
    names = argv or list(GENERATORS)
    for n in names:
        GENERATORS[n]()


if __name__ == "__main__":
    main(sys.argv[1:])
