"""CPU: the module-level entry points of vitron_b200.entrypoints keep the reference's signatures (parsed from the reference
sources when /root/reference is present, and pinned as literals for the GPU box), fail loudly when an injected host-side
object is missing, and `grounded_generation_box` reproduces the oracle chain (PLMS over the oracle UNet + scaled VAE decode)
with the kernels replaced by tests/cpu_ops_emulator.py."""
import ast
import inspect
import os
from functools import partial

import pytest
import torch

REF = "/root/reference"
EXPECTED = {
    "grounded_generation_box": (["loaded_model_list", "instruction"], "args", "kwargs"),
    "generate": (["task", "language_instruction", "grounding_texts", "sketch_pad", "alpha_sample", "guidance_scale", "batch_size",
                  "fix_seed", "rand_seed", "use_actual_mask", "append_grounding", "style_cond_image", "state", "inpainting_image",
                  "inpainting_mask"], None, None),
    "inference": (["image", "task"], "args", "kwargs"),
    "inference_i2vgen_entrance": (["cfg_update"], None, "kwargs"),
    "image_to_video": (["image_path", "text_prompt"], None, None),
    "load_pretrained_model": (["model_path", "model_base", "model_name", "load_8bit", "load_4bit", "device_map", "device"], None, "kwargs"),
}
SOURCES = {
    "grounded_generation_box": "modules/GLIGEN/demo/gligen/task_grounded_generation.py",
    "generate": "modules/GLIGEN/demo/app.py",
    "inference": "modules/SEEM/demo_code/app.py",
    "inference_i2vgen_entrance": "modules/i2vgen-xl/tools/inferences/inference_i2vgen_entrance.py",
    "image_to_video": "app.py",
    "load_pretrained_model": "vitron/model/builder.py",
}


def _sig(fn):
    s = inspect.signature(fn)
    pos = [n for n, p in s.parameters.items() if p.kind in (p.POSITIONAL_OR_KEYWORD, p.POSITIONAL_ONLY)]
    va = next((n for n, p in s.parameters.items() if p.kind == p.VAR_POSITIONAL), None)
    vk = next((n for n, p in s.parameters.items() if p.kind == p.VAR_KEYWORD), None)
    return pos, va, vk


def _ours(name):
    if name == "load_pretrained_model":
        from vitron_b200.builder import load_pretrained_model
        return load_pretrained_model
    from vitron_b200 import entrypoints
    return getattr(entrypoints, name)


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_signature_matches_reference(name):
    pos, va, vk = _sig(_ours(name))
    e_pos, e_va, e_vk = EXPECTED[name]
    assert pos[:len(e_pos)] == e_pos and va == e_va and vk == e_vk, (name, pos, va, vk)
    path = os.path.join(REF, SOURCES[name])
    if os.path.exists(path):   # the literal above is itself checked against the reference's source
        tree = ast.parse(open(path).read())
        fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == name)
        r_pos = [a.arg for a in fn.args.args]
        assert r_pos == e_pos, (name, r_pos)
        assert (fn.args.vararg.arg if fn.args.vararg else None) == e_va
        assert (fn.args.kwarg.arg if fn.args.kwarg else None) == e_vk


def test_missing_injected_objects_fail_loudly():
    from vitron_b200 import entrypoints as E
    E._STATE.clear()
    with pytest.raises(RuntimeError, match="seem_model"):
        E.inference({"image": torch.zeros((8, 8, 3), dtype=torch.uint8)}, [])
    with pytest.raises(NotImplementedError):
        E.inference({"image": torch.zeros((8, 8, 3), dtype=torch.uint8)}, ["Example"])
    with pytest.raises(RuntimeError, match="i2vgen_pipeline"):
        E.image_to_video("a.png", "a cat")
    assert E.image_to_video(None, "x") == (None, None)          # app.py:322-323
    with pytest.raises(ValueError, match="mismatching"):
        E.generate("Grounded Generation", "a dog", "dog;cat", None, 0.3, 7.5, 1, True, 0, False, True, None, {"boxes": [[0, 0, 256, 256]]})
    with pytest.raises(ValueError, match="nothing to generate"):
        E.generate("Grounded Generation", "", "", None, 0.3, 7.5, 1, True, 0, False, True, None, {"boxes": []})


def test_grounded_generation_box_host_logic(monkeypatch):
    from oracle import restate_gligen_unet as G, restate_vae as V
    from oracle.weights import seeded_state_dict
    from tests import cpu_ops_emulator
    from vitron_b200 import entrypoints as E, gligen_sampler as GS
    from vitron_b200.autoencoder import AutoencoderKL
    from vitron_b200.gligen_unet import UNetModel
    cpu_ops_emulator.install(monkeypatch)
    gold = os.path.join(os.path.dirname(__file__), "golden")
    ufx = torch.load(os.path.join(gold, "gligen_unet_tiny.pt"), weights_only=False)
    vfx = torch.load(os.path.join(gold, "vae_tiny.pt"), weights_only=False)
    usd = seeded_state_dict(ufx["shapes"], ufx["seed"], ufx["gain"])
    vsd = seeded_state_dict(vfx["shapes"], vfx["seed"], 0.8)
    cfg = dict(ufx["cfg"], image_size=8)
    unet = UNetModel(**cfg, device="cpu").load_state_dict(usd)
    vae = AutoencoderKL(vfx["ddconfig"], 4, device="cpu").load_state_dict(vsd)
    inp = ufx["inputs"]
    ctx_dim, B = inp["context"].shape[-1], 2
    g = torch.Generator().manual_seed(4)
    table = {"a dog; dog": torch.randn((1, inp["context"].shape[1], ctx_dim), generator=g), "": torch.randn((1, inp["context"].shape[1], ctx_dim), generator=g)}

    class TextEncoder:
        def encode(self, texts):
            return torch.cat([table[t] for t in texts], 0)
    pos_len = inp["text_embeddings"].shape[-1]          # 768 in the shipped GLIGEN config; narrower in the tiny fixture
    feat = torch.randn((1, pos_len), generator=g)
    clip = {"text_feature": lambda p: feat, "image_feature": lambda im: feat}
    instruction = dict(prompt="a dog; dog", phrases=["dog"], images=[None], locations=[[0.1, 0.2, 0.6, 0.7]], alpha_type=[0.4, 0.2, 0.4],
                       has_text_mask=1, has_image_mask=0, guidance_scale=2.0, batch_size=B, fix_seed=True, rand_seed=7)
    samples, overlays = E.grounded_generation_box((unet, vae, TextEncoder(), GS.DDPM()), instruction, clip_model=clip, steps=4)
    assert len(samples) == B and samples[0].dtype == torch.uint8 and samples[0].shape[-1] == 3 and overlays[0] == [[0.1, 0.2, 0.6, 0.7]]
    # oracle chain with the same seed / inputs
    torch.manual_seed(7)
    batch = E.prepare_grounding_batch(instruction, B, clip_model=clip, device="cpu", embed_dim=pos_len)
    oin = dict(x=None, timesteps=None, context=TextEncoder().encode(["a dog; dog"] * B), **{k: batch[k] for k in batch})

    class OracleModel:
        scale = 1.0
        in_channels, image_size = unet.in_channels, unet.image_size

        def __call__(self, d):
            return G.unet_forward(usd, cfg, d, alpha_scale=self.scale)
    sampler = GS.PLMSSampler(GS.DDPM(), OracleModel(), alpha_generator_func=partial(GS.alpha_generator, type=[0.4, 0.2, 0.4]),
                             set_alpha_scale=lambda m, a: setattr(m, "scale", float(a)))
    lat = sampler.sample(S=4, shape=(B, unet.in_channels, 8, 8), input=oin, uc=TextEncoder().encode([""] * B), guidance_scale=2.0)
    ref = torch.clamp(V.decode(vsd, lat / 0.18215, vfx["ddconfig"]), -1, 1) * 0.5 + 0.5
    got = torch.stack([s.permute(2, 0, 1).float() / 255 for s in samples])
    assert (got - ref).abs().max().item() < 0.12, (got - ref).abs().max().item()
