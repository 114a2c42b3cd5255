"""CPU host-logic test of the reference-callable Python face (SURVEY.md §8b face 1):
`load_pretrained_model` with the reference's signature on a synthetic LoRA checkpoint directory laid out like
`checkpoints/Vitron-base` + `checkpoints/Vitron-lora` (builder.py:52-84), the nn.Module surface of the drop-in
(`state_dict()/.to()/.eval()/.parameters()`), and then the call sequence of `inference_image.py:19-61` line for line
(processor['image'].preprocess -> model.generate(input_ids, images=..., do_sample..., stopping_criteria=[...]) ->
tokenizer.decode). Kernels are replaced by the torch statements of tests/cpu_ops_emulator.py; the generated ids must
equal the CPU oracle's greedy ids on the MERGED weights."""
import json
import os

import pytest
import torch

from oracle import restate_llm as R
from oracle.weights import seeded_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")


class DummyTokenizer:
    """Stands in for the LLaMA sentencepiece tokenizer (no tokenizer files in the tree, checkpoints/README.md:15)."""
    padding_side, model_max_length = "right", 4096

    def __init__(self, n):
        self.n, self.added = n, []

    def add_tokens(self, toks, special_tokens=False):
        self.added += [t for t in toks if t not in self.added]
        return len(toks)

    def __len__(self):
        return self.n + len(self.added)

    def decode(self, ids):
        return " ".join(str(int(i)) for i in ids)


class KeywordsStoppingCriteria:
    """vitron/mm_utils.py KeywordsStoppingCriteria reduced to ids: stop when the keyword id was generated."""

    def __init__(self, keyword_id, input_ids):
        self.keyword_id, self.start_len = keyword_id, input_ids.shape[1]

    def __call__(self, output_ids, scores=None, **kw):
        return bool((output_ids[:, self.start_len:] == self.keyword_id).any())


@pytest.fixture()
def checkpoint(tmp_path):
    fx = torch.load(os.path.join(GOLD, "vitron_llm_tiny.pt"), weights_only=False)
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    llm, vit = fx["llm"], fx["vit"]
    base, lora, cache = tmp_path / "Vitron-base", tmp_path / "Vitron-lora", tmp_path / "cache_dir"
    for d in (base, lora, cache / "LanguageBind_Image", cache / "LanguageBind_Video_merge"):
        d.mkdir(parents=True)
    cfg = dict(llm, model_type="llava", mm_image_tower="LanguageBind_Image", mm_video_tower="LanguageBind_Video_merge",
               mm_projector_type="mlp2x_gelu", mm_vision_select_layer=-2, mm_vision_select_feature="patch",
               mm_use_im_start_end=False, mm_use_im_patch_token=True, rms_norm_eps=1e-5, rope_theta=10000.0)
    json.dump(cfg, open(lora / "config.json", "w"))
    json.dump(dict(vision_config=dict(vit, hidden_act="gelu")), open(cache / "LanguageBind_Image" / "config.json", "w"))
    json.dump(dict(vision_config=dict(vit, hidden_act="gelu", num_frames=fx["num_frames"])),
              open(cache / "LanguageBind_Video_merge" / "config.json", "w"))
    # base checkpoint = everything except the projector / region extractor, split over two shards with an index
    adapters = {k: v for k, v in sd.items() if k.startswith(("model.mm_projector.", "model.region_extractor."))}
    rest = {k: v.to(torch.bfloat16) for k, v in sd.items() if k not in adapters}
    names = sorted(rest)
    shards = {"pytorch_model-00001-of-00002.bin": names[:len(names) // 2], "pytorch_model-00002-of-00002.bin": names[len(names) // 2:]}
    for fn, ks in shards.items():
        torch.save({k: rest[k] for k in ks}, base / fn)
    json.dump({"weight_map": {k: fn for fn, ks in shards.items() for k in ks}}, open(base / "pytorch_model.bin.index.json", "w"))
    # LoRA directory: non_lora_trainables.bin with the reference's key prefixes + a rank-2 adapter on two projections
    torch.save({"base_model.model." + k: v for k, v in adapters.items()}, lora / "non_lora_trainables.bin")
    g = torch.Generator().manual_seed(5)
    d = llm["hidden_size"]
    ad, merged = {}, dict(sd)
    for tgt in ("model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj"):
        w = sd[tgt + ".weight"]
        A, B = torch.randn((2, w.shape[1]), generator=g) * 0.05, torch.randn((w.shape[0], 2), generator=g) * 0.05
        ad[f"base_model.model.{tgt}.lora_A.weight"], ad[f"base_model.model.{tgt}.lora_B.weight"] = A, B
        merged[tgt + ".weight"] = (w.to(torch.bfloat16).float() + (B @ A) * (8 / 2)).to(torch.bfloat16).float()
    torch.save(ad, lora / "adapter_model.bin")
    json.dump(dict(r=2, lora_alpha=8, target_modules=["q_proj", "down_proj"]), open(lora / "adapter_config.json", "w"))
    return fx, merged, str(base), str(lora), str(cache)


def test_load_pretrained_model_and_inference_image_sequence(checkpoint, monkeypatch):
    from tests import cpu_ops_emulator
    from tests.test_oracle_cpu import cfgs_of
    cpu_ops_emulator.install(monkeypatch)
    from vitron_b200.builder import load_pretrained_model
    fx, merged, model_base, model_path, cache_dir = checkpoint
    V = fx["llm"]["vocab_size"]
    with pytest.raises(ValueError):
        load_pretrained_model(model_path, model_base, "vitron-llava-7b-lora-4", True, False, device="cpu", tokenizer=DummyTokenizer(V))
    # ---- inference_image.py:19
    load_4bit, load_8bit = False, False
    tokenizer, model, processor, context_len = load_pretrained_model(
        model_path, model_base, "vitron-llava-7b-lora-4", load_8bit, load_4bit, device="cpu", cache_dir=cache_dir,
        tokenizer=DummyTokenizer(V), max_batch=2, max_seq_len=256)
    assert context_len == 2048 and set(processor) == {"image", "video"} and processor["image"] is not None and processor["video"] is not None
    assert len(tokenizer) == V + 2 and model.config.vocab_size == V + 2           # <im_patch>, <vid_patch> (builder.py:139-146)
    # ---- nn.Module face
    assert model.eval() is model and model.to("cpu") is model and model.to(dtype=torch.float16) is model
    with pytest.raises(ValueError):
        model.to("cuda:3")
    assert next(iter(model.parameters())).dtype == torch.bfloat16 and str(model.device) == "cpu"
    got = model.state_dict()
    for k, v in merged.items():
        ref = v.to(torch.bfloat16).float()
        g_ = got[k].float()
        if k in ("model.embed_tokens.weight", "lm_head.weight"):
            assert g_.shape[0] == V + 2 and bool((g_[V:] == 0).all())
            g_ = g_[:V]
        assert g_.shape == ref.shape, k
        # folded RMSNorm weights come back within one bf16 rounding; everything else exactly
        tol = 2 ** -7 * ref.abs().max().item()
        assert (g_ - ref).abs().max().item() <= tol, (k, (g_ - ref).abs().max().item())
    # ---- inference_image.py:20-61, line for line (PIL / conversation-template lines replaced by their products)
    image_processor = processor["image"]
    raw = torch.randint(0, 256, (40, 56, 3), generator=torch.Generator().manual_seed(2), dtype=torch.uint8)  # Image.open(image)
    vit_size = fx["vit"]["image_size"]
    if vit_size == 224:
        image_tensor = image_processor.preprocess(raw, return_tensors="pt")["pixel_values"]
    else:   # the tiny tower of the fixture is not 224 x 224: feed the fixture's own pixels through the same code path
        image_tensor = torch.stack(fx["gen_img"]["images"])[:1]
    if type(image_tensor) is list:
        tensor = [image.to(model.device, dtype=torch.float16) for image in image_tensor]
    else:
        tensor = image_tensor.to(model.device, dtype=torch.float16)
    input_ids = fx["gen_img"]["input_ids"][:1].clone()                          # tokenizer_image_region_token(...).unsqueeze(0)
    ids_no_region = input_ids[input_ids != -300].view(1, -1)                    # `# regions = region,` is commented out upstream
    c = cfgs_of(fx)
    want, gaps = R.greedy_generate({k: v.to(torch.bfloat16).float() for k, v in merged.items()}, c, ids_no_region,
                                   [t for t in tensor.float()], None, 6)
    stopping_criteria = KeywordsStoppingCriteria(int(want[0, 3]), ids_no_region)
    with torch.inference_mode():
        output_ids = model.generate(
            ids_no_region,
            images=tensor,
            do_sample=False,
            temperature=0.2,
            max_new_tokens=1024 if False else 6,
            use_cache=True,
            stopping_criteria=[stopping_criteria])
    outputs = tokenizer.decode(output_ids[0, ids_no_region.shape[1]:]).strip()
    new = output_ids[0, ids_no_region.shape[1]:]
    assert len(outputs) > 0 and 1 <= new.shape[0] <= 6
    tol = 2 * 0.04 * fx["img"]["logits"].abs().max().item()
    n_cmp = 0
    for t in range(new.shape[0]):
        if gaps[0, t] <= tol:
            break
        assert int(new[t]) == int(want[0, t]), (new.tolist(), want.tolist())
        n_cmp += 1
    assert n_cmp >= 1
    # the stopping criterion fires on the keyword: nothing is returned beyond it
    if n_cmp >= 4:
        assert int(new[-1]) == int(want[0, 3]) and new.shape[0] == 4
    # sampled decoding with the reference's keywords runs through the same engine
    torch.manual_seed(0)
    with torch.inference_mode():
        out2 = model.generate(ids_no_region, images=tensor, do_sample=True, temperature=0.2, max_new_tokens=4, use_cache=True)
    assert out2.shape[1] <= ids_no_region.shape[1] + 4 and bool((out2[:, :ids_no_region.shape[1]] == ids_no_region).all())
