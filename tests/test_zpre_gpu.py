"""GPU parity for SURVEY.md §8(f3): the fused device-side LanguageBind image / video transform (preprocess.cu) against
the CPU restatement built from the same ATen interpolation operators the reference's torchvision / pytorchvideo
transforms call (oracle/restate_preprocess.py; pinned against the reference's own get_image_transform on CPU).
fp32 outputs: |err| <= 2e-4 in normalised units (the resample is a <= 400-tap fp32 dot product of uint8 values)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def u8(shape, seed):
    return torch.randint(0, 256, shape, generator=torch.Generator().manual_seed(seed), dtype=torch.uint8)


@pytest.mark.parametrize("h,w", [(336, 336), (224, 224), (480, 640), (500, 333), (100, 150), (1024, 1024), (225, 224)])
@pytest.mark.parametrize("antialias", [False, True])
def test_image_transform(cuda, h, w, antialias):
    from oracle import restate_preprocess as P
    from vitron_b200.processing import LanguageBindImageProcessor
    img = u8((h, w, 3), h * 7 + w)
    proc = LanguageBindImageProcessor(device=cuda, antialias=antialias)
    got = proc.preprocess(img.numpy())["pixel_values"]
    ref = P.image_transform(img, antialias)
    assert got.shape == (1, 3, 224, 224) and got.dtype == torch.float32
    err = (got[0].cpu() - ref).abs().max().item()
    assert err <= 2e-4, err
    b = LanguageBindImageProcessor(device=cuda, antialias=antialias, dtype=torch.bfloat16).preprocess([img.numpy(), img.numpy()])
    assert b["pixel_values"].shape == (2, 3, 224, 224)
    assert (b["pixel_values"][1].float().cpu() - ref).abs().max().item() <= 2e-2


@pytest.mark.parametrize("h,w,flip", [(240, 320, False), (360, 640, True), (300, 225, True), (224, 224, False), (720, 1280, False)])
def test_video_transform(cuda, h, w, flip):
    from oracle import restate_preprocess as P
    from vitron_b200.processing import LanguageBindVideoProcessor, sample_frame_ids
    clip = u8((20, h, w, 3), h + w)
    ids = sample_frame_ids(20, 8)
    frames = clip[torch.from_numpy(ids)]
    got = LanguageBindVideoProcessor(device=cuda).preprocess(frames, flip=flip)["pixel_values"]
    ref = P.video_transform(frames, flip)
    assert got.shape == (1, 3, 8, 224, 224)
    err = (got[0].cpu() - ref).abs().max().item()
    assert err <= 2e-4, err


def test_processed_pixels_feed_the_tower_entry(cuda):
    """The processor output has the layout / dtype `encode_images` takes ([B, 3, 224, 224] float on the device)."""
    from vitron_b200.processing import LanguageBindImageProcessor
    out = LanguageBindImageProcessor(device=cuda).preprocess([u8((336, 336, 3), 1).numpy()] * 3)["pixel_values"]
    assert out.is_cuda and out.is_contiguous() and tuple(out.shape) == (3, 3, 224, 224)
    batch = torch.stack([u8((336, 336, 3), 1)] * 3)
    one = LanguageBindImageProcessor(device=cuda).preprocess(batch)["pixel_values"]
    assert torch.equal(one, out), "batched launch == per-image launches"
    with pytest.raises(ValueError):
        LanguageBindImageProcessor(device=cuda)(images=None, text=None)
