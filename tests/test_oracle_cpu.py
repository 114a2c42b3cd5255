"""CPU: pins the oracle restatements (oracle/restate_*.py) against golden vectors produced by the
unmodified reference (oracle/gen_golden.py) and, when /root/reference is present, against the
reference itself on a fresh random configuration."""
import os

import pytest
import torch

from oracle import restate_llm as R
from oracle.weights import seeded_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def cfgs_of(fx):
    vit = dict(fx["vit"], hidden_act="gelu", layer_norm_eps=1e-5)
    return {"llm": dict(fx["llm"], rms_norm_eps=1e-5, rope_theta=10000.0), "vision": dict(vit, add_time_attn=False),
            "video": dict(vit, add_time_attn=True, num_frames=fx["num_frames"]), "max_len": 4096}


@pytest.fixture(scope="module")
def fx():
    return load("vitron_llm_tiny.pt")


@pytest.fixture(scope="module")
def sd(fx):
    return seeded_state_dict(fx["shapes"], fx["seed"])


def test_towers_and_adapters_match_reference_golden(fx, sd):
    c = cfgs_of(fx)
    t = fx["tower"]
    imgs = torch.stack(fx["img"]["images"])
    feats = R.clip_vit_hidden(sd, "model.image_tower.image_tower.", c["vision"], imgs)[:, 1:]
    assert torch.allclose(feats, t["image_feats"], atol=2e-4, rtol=1e-4)
    vf = R.clip_vit_hidden(sd, "model.video_tower.video_tower.", c["video"], fx["vid"]["images"][0][None])[:, :, 1:]
    assert torch.allclose(vf, t["video_feats"], atol=2e-4, rtol=1e-4)
    assert torch.allclose(R.projector(sd, "model.mm_projector.", feats), t["proj"], atol=2e-4, rtol=1e-4)
    rg = R.region_extractor(sd, "model.region_extractor.", feats, fx["img"]["regions"], fx["vit"]["image_size"])
    assert torch.allclose(rg, t["region"], atol=2e-4, rtol=1e-4)


def test_multimodal_logits_match_reference_golden(fx, sd):
    c = cfgs_of(fx)
    g = fx["img"]
    logits, lens = R.vitron_logits(sd, c, g["input_ids"], g["images"], g["regions"], g["attention_mask"])
    assert logits.shape == g["logits"].shape
    for b, n in enumerate(lens):  # padded positions are don't-care
        assert torch.allclose(logits[b, :n], g["logits"][b, :n], atol=2e-3, rtol=1e-3), b
    v = fx["vid"]
    lv, _ = R.vitron_logits(sd, c, v["input_ids"], v["images"])
    assert torch.allclose(lv, v["logits"], atol=2e-3, rtol=1e-3)


def test_greedy_tokens_match_reference_golden(fx, sd):
    c = cfgs_of(fx)
    g = fx["gen_img"]
    toks, gaps = R.greedy_generate(sd, c, g["input_ids"], g["images"], g["regions"], g["tokens"].shape[1])
    assert torch.equal(toks, g["tokens"]), (toks, g["tokens"], gaps)
    v = fx["vid"]
    toks, _ = R.greedy_generate(sd, c, v["input_ids"], v["images"], None, v["tokens"].shape[1])
    assert torch.equal(toks, v["tokens"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/vitron"), reason="reference tree not present")
def test_oracle_against_live_reference():
    from oracle import refshim
    from oracle.weights import shapes_of
    llm = dict(hidden_size=64, intermediate_size=96, num_hidden_layers=3, num_attention_heads=4, vocab_size=97)
    vit = dict(hidden_size=48, intermediate_size=80, num_hidden_layers=4, num_attention_heads=3, image_size=42, patch_size=14)
    model = refshim.build_reference_vitron(llm, vit, with_video=False, hidden_act="quick_gelu")
    sd = seeded_state_dict(shapes_of(model), 5)
    model.load_state_dict(sd, strict=False)
    g = torch.Generator().manual_seed(1)
    img = torch.randn((3, 42, 42), generator=g)
    ids = torch.tensor([[1, 4, 9, -200, 17, 33, -300, 8]])
    with torch.no_grad():
        ref = model(input_ids=ids, images=[img], regions=[[3.0, 5.0, 30.0, 40.0]], use_cache=False).logits
    c = {"llm": dict(llm, rms_norm_eps=1e-5, rope_theta=10000.0),
         "vision": dict(vit, hidden_act="quick_gelu", layer_norm_eps=1e-5, add_time_attn=False), "max_len": 4096}
    got, _ = R.vitron_logits(sd, c, ids, [img], [[3.0, 5.0, 30.0, 40.0]])
    assert torch.allclose(got, ref.float(), atol=2e-3, rtol=1e-3)


def test_cached_decoder_equals_full_recompute(fx, sd):
    c = cfgs_of(fx)
    sdf = {k: v.float() for k, v in sd.items()}
    g = torch.Generator().manual_seed(0)
    emb = torch.randn((2, 9, fx["llm"]["hidden_size"]), generator=g)
    m = R.LlamaCPU(sdf, c["llm"])
    lg = m.prefill(emb)
    full = R.llama_forward(sdf, c["llm"], emb)
    assert torch.allclose(lg, full[:, -1], atol=1e-4, rtol=1e-4)
    tok = lg.argmax(-1)
    lg2 = m.step(tok)
    emb2 = torch.cat([emb, sdf["model.embed_tokens.weight"][tok][:, None]], 1)
    assert torch.allclose(lg2, R.llama_forward(sdf, c["llm"], emb2)[:, -1], atol=1e-4, rtol=1e-4)


# ------------------------------------------------------------------ i2vgen-xl UNet3D + DDIM
@pytest.fixture(scope="module")
def ufx():
    return load("unet_tiny.pt")


def test_unet_restatement_matches_reference_golden(ufx):
    from oracle import restate_unet as U
    sd = seeded_state_dict(ufx["shapes"], ufx["seed"], ufx["gain"])
    i = ufx["inputs"]
    out = U.unet_forward(sd, ufx["cfg"], i["x"], i["t"], y=i["y"], image=i["image"], local_image=i["local_image"], fps=i["fps"])
    ref = ufx["out"]
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, atol=1e-3 * ref.abs().max().item(), rtol=1e-3), (out - ref).abs().max()


def test_ddim_restatement_matches_reference_golden(ufx):
    from oracle import restate_unet as U
    sd = seeded_state_dict(ufx["shapes"], ufx["seed"], ufx["gain"])
    d = ufx["ddim"]
    model = lambda xt, t, **kw: U.unet_forward(sd, ufx["cfg"], xt, t, **kw)
    out = U.ddim_sample_loop(d["noise"], model, [d["cond"], d["uncond"]], d["guide_scale"], d["ddim_timesteps"])
    assert torch.allclose(out, d["out"], atol=2e-3, rtol=2e-3), (out - d["out"]).abs().max()


# ------------------------------------------------------------------ GLIGEN gated self-attention
def test_gligen_restatement_matches_reference_golden():
    from oracle import restate_gligen as G
    fx = load("gligen_tiny.pt")
    for c in fx["cases"]:
        sd = seeded_state_dict(c["shapes"], fx["seed"])
        fo = G.gated_self_attention_dense(sd, "fuser.", c["x"], fx["objs"], c["heads"])
        assert torch.allclose(fo, c["fuser_out"], atol=1e-4, rtol=1e-4)
        bo = G.basic_transformer_block(sd, "", c["x"], fx["context"], fx["objs"], c["heads"])
        assert torch.allclose(bo, c["block_out"], atol=1e-4, rtol=1e-4)
