"""GPU parity of the vision-LLM path against (i) golden vectors from the
unmodified reference and (ii) the CPU oracle at a mid-size configuration. bf16 kernels vs fp32
reference: tolerance = 4% of the largest reference magnitude (inf-norm) and 3% relative L2;
greedy token ids must be identical wherever the oracle's top-2 logit margin exceeds that tolerance."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def assert_close(got, ref, what, rel_inf=0.04, rel_l2=0.03):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = ref.abs().max().item() + 1e-6
    e_inf = (got - ref).abs().max().item() / scale
    e_l2 = ((got - ref).norm() / (ref.norm() + 1e-6)).item()
    assert e_inf < rel_inf and e_l2 < rel_l2, f"{what}: inf {e_inf:.4f} l2 {e_l2:.4f}"


def build_model(cuda, llm, vit, num_frames, sd, hidden_act="gelu", max_seq_len=512, max_batch=4):
    from vitron_b200.vision_tower import VisionConfig
    from vitron_b200.vitron_model import VitronConfig, VitronLlamaForCausalLM
    vis = VisionConfig(**vit, hidden_act=hidden_act)
    vid = VisionConfig(**vit, hidden_act=hidden_act, add_time_attn=True, num_frames=num_frames) if num_frames else None
    cfg = VitronConfig(llm=llm, vision=vis, video=vid, tokenizer_model_max_length=4096)
    m = VitronLlamaForCausalLM(cfg, cuda, max_batch=max_batch, max_seq_len=max_seq_len)
    m.load_state_dict(sd)
    return m


@pytest.fixture(scope="module")
def golden(cuda):
    from oracle.weights import seeded_state_dict
    fx = torch.load(os.path.join(GOLD, "vitron_llm_tiny.pt"), weights_only=False)
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    model = build_model(cuda, fx["llm"], fx["vit"], fx["num_frames"], sd)
    return fx, sd, model


def test_towers_and_adapters_vs_reference_golden(cuda, golden):
    fx, sd, m = golden
    t = fx["tower"]
    imgs = torch.stack(fx["img"]["images"]).to(cuda)
    feats = m.get_image_tower()(imgs)
    assert_close(feats, t["image_feats"], "image tower")
    vf = m.get_video_tower()(fx["vid"]["images"][0][None].to(cuda))
    assert_close(vf, t["video_feats"], "video tower")
    assert_close(m.get_model().mm_projector(feats), t["proj"], "projector")
    assert_close(m.get_region_extractor()(feats, fx["img"]["regions"]), t["region"], "region extractor")


def test_multimodal_logits_vs_reference_golden(cuda, golden):
    fx, sd, m = golden
    g = fx["img"]
    out = m.forward(input_ids=g["input_ids"].to(cuda), attention_mask=g["attention_mask"].to(cuda),
                    images=[i.to(cuda) for i in g["images"]], regions=g["regions"])
    lens = m._last_lens
    for b, n in enumerate(lens):
        assert_close(out.logits[b, :n], g["logits"][b, :n], f"logits sample {b}")
    v = fx["vid"]
    out = m.forward(input_ids=v["input_ids"].to(cuda), images=[v["images"][0].to(cuda)])
    assert_close(out.logits, v["logits"], "video logits")


def _check_tokens(got, ref, gaps, tol):
    """identical until the first step whose oracle margin is inside the tolerance; returns how many tokens were compared (the
    callers assert a minimum: on these tiny random models the margins are small, so the count is what the oracle allows — the
    teacher-forced 7B test in test_fullsize_gpu.py compares every decode position)."""
    compared = 0
    for b in range(ref.shape[0]):
        for t in range(ref.shape[1]):
            if gaps[b, t] <= tol:
                break
            assert int(got[b, t]) == int(ref[b, t]), f"token mismatch b={b} t={t}: {got[b].tolist()} vs {ref[b].tolist()} (margin {gaps[b, t]:.3f})"
            compared += 1
    return compared


def test_greedy_generate_vs_reference_golden(cuda, golden):
    from oracle import restate_llm as R
    fx, sd, m = golden
    from tests.test_oracle_cpu import cfgs_of
    c = cfgs_of(fx)
    for key in ("gen_img", "vid"):
        g = fx[key]
        n = g["tokens"].shape[1]
        reg = g.get("regions")
        otoks, gaps = R.greedy_generate(sd, c, g["input_ids"], g["images"], reg, n)
        assert torch.equal(otoks, g["tokens"])  # oracle == reference (also checked on CPU)
        out = m.generate(g["input_ids"].to(cuda), images=[i.to(cuda) for i in g["images"]], regions=reg,
                         do_sample=False, max_new_tokens=n, use_cache=True, eos_token_id=-1)
        got = out[:, g["input_ids"].shape[1]:].cpu()
        assert got.shape == g["tokens"].shape
        scale = 0.04 * fx[("img" if key == "gen_img" else "vid")]["logits"].abs().max().item()
        assert _check_tokens(got, g["tokens"], gaps, 2 * scale) >= (2 if key == "gen_img" else 1)   # oracle-decisive prefix of this fixture


def test_midsize_vs_oracle(cuda):
    """LLaMA hd=128 x4 heads, ViT hd=64, 64 patches, ragged batch of 3 with regions; graph decode."""
    from oracle import restate_llm as R
    from oracle.weights import seeded_state_dict
    llm = dict(hidden_size=512, intermediate_size=1408, num_hidden_layers=4, num_attention_heads=4, vocab_size=2000)
    vit = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=4, image_size=112, patch_size=14)
    shapes = {}
    d, f, V = llm["hidden_size"], llm["intermediate_size"], llm["vocab_size"]
    shapes["model.embed_tokens.weight"] = [V, d]; shapes["lm_head.weight"] = [V, d]; shapes["model.norm.weight"] = [d]
    for i in range(llm["num_hidden_layers"]):
        p = f"model.layers.{i}."
        for n in "qkvo":
            shapes[p + f"self_attn.{n}_proj.weight"] = [d, d]
        shapes[p + "mlp.gate_proj.weight"] = [f, d]; shapes[p + "mlp.up_proj.weight"] = [f, d]; shapes[p + "mlp.down_proj.weight"] = [d, f]
        shapes[p + "input_layernorm.weight"] = [d]; shapes[p + "post_attention_layernorm.weight"] = [d]
    vd, vf = vit["hidden_size"], vit["intermediate_size"]
    npatch = (vit["image_size"] // 14) ** 2
    vp = "model.image_tower.image_tower."
    shapes[vp + "embeddings.class_embedding"] = [vd]; shapes[vp + "embeddings.patch_embedding.weight"] = [vd, 3, 14, 14]
    shapes[vp + "embeddings.position_embedding.weight"] = [npatch + 1, vd]
    shapes[vp + "pre_layrnorm.weight"] = [vd]; shapes[vp + "pre_layrnorm.bias"] = [vd]
    for i in range(vit["num_hidden_layers"]):
        p = vp + f"encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            shapes[p + f"self_attn.{n}.weight"] = [vd, vd]; shapes[p + f"self_attn.{n}.bias"] = [vd]
        for n in ("layer_norm1", "layer_norm2"):
            shapes[p + n + ".weight"] = [vd]; shapes[p + n + ".bias"] = [vd]
        shapes[p + "mlp.fc1.weight"] = [vf, vd]; shapes[p + "mlp.fc1.bias"] = [vf]
        shapes[p + "mlp.fc2.weight"] = [vd, vf]; shapes[p + "mlp.fc2.bias"] = [vd]
    shapes["model.mm_projector.0.weight"] = [d, vd]; shapes["model.mm_projector.0.bias"] = [d]
    shapes["model.mm_projector.2.weight"] = [d, d]; shapes["model.mm_projector.2.bias"] = [d]
    rp = "model.region_extractor."
    for i, (o, k) in enumerate([(d, vd), (d, d), (d, d)]):
        shapes[rp + f"region_linear.layers.{i}.weight"] = [o, k]; shapes[rp + f"region_linear.layers.{i}.bias"] = [o]
    shapes[rp + "loc_encoder.loc_encoder.0.weight"] = [d // 2, 4]; shapes[rp + "loc_encoder.loc_encoder.0.bias"] = [d // 2]
    shapes[rp + "loc_encoder.loc_encoder.2.weight"] = [d, d // 2]; shapes[rp + "loc_encoder.loc_encoder.2.bias"] = [d]
    sd = seeded_state_dict(shapes, 3)
    m = build_model(cuda, llm, vit, 0, sd, hidden_act="quick_gelu", max_seq_len=256, max_batch=4)
    g = torch.Generator().manual_seed(9)
    imgs = [torch.randn((3, 112, 112), generator=g) for _ in range(3)]
    regions = [[10.0, 20.0, 90.0, 100.0], [0.0, 0.0, 112.0, 60.0], [50.0, 50.0, 51.0, 112.0]]
    rows = [[1] + torch.randint(3, V, (n0,), generator=g).tolist() + [-200] + torch.randint(3, V, (n1,), generator=g).tolist()
            + ([-300] if objs else []) + torch.randint(3, V, (2,), generator=g).tolist()
            for n0, n1, objs in ((4, 9, True), (1, 3, False), (7, 20, True))]
    L = max(len(r) for r in rows)
    ids = torch.zeros((3, L), dtype=torch.long); am = torch.zeros((3, L), dtype=torch.long)
    for b, r in enumerate(rows):
        ids[b, :len(r)] = torch.tensor(r); am[b, :len(r)] = 1
    c = {"llm": dict(llm, rms_norm_eps=1e-5, rope_theta=10000.0),
         "vision": dict(vit, hidden_act="quick_gelu", layer_norm_eps=1e-5, add_time_attn=False), "max_len": 4096}
    ref, lens = R.vitron_logits(sd, c, ids, imgs, regions, am)
    out = m.forward(input_ids=ids.to(cuda), attention_mask=am.to(cuda), images=[i.to(cuda) for i in imgs], regions=regions)
    assert m._last_lens == lens
    for b, n in enumerate(lens):
        assert_close(out.logits[b, :n], ref[b, :n], f"midsize logits {b}")
    n_new = 20
    otoks, gaps = R.greedy_generate(sd, c, ids, imgs, regions, n_new, am)
    tol = 2 * 0.04 * ref.abs().max().item()
    for use_graph_chunk in (16, 3):
        got = m.generate(ids.to(cuda), attention_mask=am.to(cuda), images=[i.to(cuda) for i in imgs], regions=regions,
                         do_sample=False, max_new_tokens=n_new, eos_token_id=-1, sync_every=use_graph_chunk)
        assert _check_tokens(got[:, L:].cpu(), otoks, gaps, tol) >= 2   # sample 2's first two steps are oracle-decisive
    # eos handling: force eos = first generated token of sample 0 -> that row is padded afterwards
    eos = int(otoks[0, 0])
    got = m.generate(ids.to(cuda), attention_mask=am.to(cuda), images=[i.to(cuda) for i in imgs], regions=regions,
                     do_sample=False, max_new_tokens=6, eos_token_id=eos, pad_token_id=0)
    assert int(got[0, L]) == eos and (got[0, L + 1:] == 0).all()
