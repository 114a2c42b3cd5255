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


def _load_seeded(model, seed):
    shapes = shapes_of(model)
    sd = seeded_state_dict(shapes, seed)
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


GENERATORS = {"vitron_llm": gen_vitron_llm}


def main(argv):
    os.makedirs(OUT, exist_ok=True)
    names = argv or list(GENERATORS)
    for n in names:
        GENERATORS[n]()


if __name__ == "__main__":
    main(sys.argv[1:])
