"""GPU parity for SURVEY.md §8(f2): i2vgen-xl's OpenCLIP text / image embedders (tools/modules/clip_embedder.py) on the
vitron_b200 kernels against the CPU restatement of the published open_clip algorithm (oracle/restate_openclip.py,
cross-checked against transformers' CLIP; open_clip itself is absent -> parity unpinned against it). ViT-H-14 widths
(text 1024 x 16 heads x 77 tokens causal; vision 1280 x 16 heads of 80) at reduced depth; <= 4 % inf / 3 % L2."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def assert_close(got, ref, what, rel_inf=0.04, rel_l2=0.03):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = ref.abs().max().item() + 1e-6
    e_inf = (got - ref).abs().max().item() / scale
    e_l2 = ((got - ref).norm() / (ref.norm() + 1e-6)).item()
    assert e_inf < rel_inf and e_l2 < rel_l2, f"{what}: inf {e_inf:.4f} l2 {e_l2:.4f}"


@pytest.mark.parametrize("layer", ["penultimate", "last"])
def test_openclip_text_and_image_embedder(cuda, layer):
    from oracle import restate_openclip as OC
    from oracle.weights import seeded_state_dict
    from vitron_b200.clip_embedder import VIT_H_14, FrozenOpenCLIPTtxtVisualEmbedder
    cfg = dict(VIT_H_14, text=dict(VIT_H_14["text"], layers=4, vocab_size=2048), vision=dict(VIT_H_14["vision"], layers=3))
    sd = seeded_state_dict(OC.openclip_shapes(cfg), 17)
    g = torch.Generator().manual_seed(4)
    tokens = torch.randint(1, 2047, (3, 77), generator=g)
    for b, e in enumerate((5, 40, 76)):            # SOT ... EOT (highest id) then padding zeros, like open_clip.tokenize
        tokens[b, e] = 2047
        tokens[b, e + 1:] = 0
    img = torch.randn((2, 3, 224, 224), generator=g)
    emb = FrozenOpenCLIPTtxtVisualEmbedder(None, device=cuda, layer=layer, arch_cfg=cfg).load_state_dict(sd)
    xi, xt, x = emb(image=img.to(cuda), text=tokens.to(cuda))
    rt, rx = OC.encode_text(sd, tokens, cfg, layer_idx=1 if layer == "penultimate" else 0)
    assert_close(x, rx, "y_words (ln_final tokens)")
    assert_close(xt, rt, "y_text (EOS pooled @ text_projection)")
    assert_close(xi, OC.encode_image(sd, img, cfg), "y_visual (encode_image)")
