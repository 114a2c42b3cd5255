"""GPU parity for SURVEY.md §8(f2): GLIGEN's grounded SD UNet (openaimodel.py::UNetModel) on the vitron_b200 kernels:
`vitron_b200.gligen_unet.UNetModel` against the golden output of the UNMODIFIED reference class
(tests/golden/gligen_unet_tiny.pt) and against the pinned CPU restatement (oracle/restate_gligen_unet.py) at the SD-1.4
widths (320 channels, 8 heads of 40 / 80 / 160, context 768) incl. the inpainting input and the no-grounding path;
<= 5 % of the reference inf-norm and <= 4 % relative L2."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def assert_close(got, ref, what, rel_inf=0.05, rel_l2=0.04):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = ref.abs().max().item() + 1e-6
    e_inf = (got - ref).abs().max().item() / scale
    e_l2 = ((got - ref).norm() / (ref.norm() + 1e-6)).item()
    assert e_inf < rel_inf and e_l2 < rel_l2, f"{what}: inf {e_inf:.4f} l2 {e_l2:.4f}"


def test_gligen_unet_vs_reference_golden(cuda):
    from oracle.weights import seeded_state_dict
    from vitron_b200.gligen_unet import UNetModel
    fx = torch.load(os.path.join(GOLD, "gligen_unet_tiny.pt"), weights_only=False)
    net = UNetModel(**fx["cfg"], device=cuda).load_state_dict(seeded_state_dict(fx["shapes"], fx["seed"], fx["gain"]))
    out = net({k: v.to(cuda) for k, v in fx["inputs"].items()})
    assert_close(out, fx["out"], "golden eps prediction")


@pytest.mark.parametrize("variant", ["sd14", "inpaint_nogrounding"])
def test_gligen_unet_vs_oracle_sd_widths(cuda, variant):
    from oracle import restate_gligen_unet as G
    from oracle.weights import seeded_state_dict
    from vitron_b200 import param_shapes
    from vitron_b200.gligen_unet import SD14_GLIGEN_UNET, UNetModel
    cfg = dict(SD14_GLIGEN_UNET, image_size=16, is_inpaint=(variant != "sd14"))
    sd = seeded_state_dict(param_shapes.gligen_unet_shapes(cfg), 5, 0.6)
    g = torch.Generator().manual_seed(9)
    rn = lambda *s: torch.randn(s, generator=g)
    b, hw = (2, (16, 16)) if variant == "sd14" else (1, (16, 24))
    inp = dict(x=rn(b, 4, *hw), timesteps=torch.tensor([981, 250][:b]), context=rn(b, 77, 768))
    if variant == "sd14":
        inp.update(boxes=torch.rand((b, 30, 4), generator=g), masks=(torch.rand((b, 30), generator=g) > 0.7).float(),
                   text_embeddings=rn(b, 30, 768))
    else:
        inp["inpainting_extra_input"] = rn(b, 5, *hw)
    ref = G.unet_forward(sd, cfg, inp)
    out = UNetModel(**cfg, device=cuda).load_state_dict(sd)({k: v.to(cuda) for k, v in inp.items()})
    assert_close(out, ref, f"{variant} eps prediction")
