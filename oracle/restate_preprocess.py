"""TEST INFRASTRUCTURE ONLY — CPU restatement of the LanguageBind image / video transforms with
the same ATen interpolation operators torchvision / pytorchvideo call; never imported by the product package.

Image: vitron/model/multimodal_encoder/languagebind/image/processing_image.py:15-25 — ToTensor, Resize(224, BICUBIC)
(on a float tensor: torchvision -> F.interpolate(mode="bicubic", align_corners=False, antialias=<version default>)),
CenterCrop(224), Normalize. PINNED for the antialiased variant against the reference's own `get_image_transform` run
with the torchvision of this image (0.26: antialias=True); the non-antialiased variant is what the reference's pinned
torchvision==0.15.2 (pyproject.toml:16) does for tensor inputs and differs only in the `antialias` flag of the same call.
Video: video/processing_video.py:26-70 — /255, NormalizeVideo, ShortSideScale(224), CenterCropVideo(224), flip.
pytorchvideo is absent from this image (third-party, requirements pin pytorchvideo==0.1.5): ShortSideScale is restated
from its published definition (pytorchvideo/transforms/functional.py::short_side_scale: floor(long/short*size),
F.interpolate(mode="bilinear", align_corners=False)); CenterCropVideo / NormalizeVideo are torchvision's
_transforms_video (crop offsets int(round((h - th) / 2.))). Parity unpinned for the video variant beyond that.
"""
import math

import torch
import torch.nn.functional as F

MEAN = (0.48145466, 0.4578275, 0.40821073)
STD = (0.26862954, 0.26130258, 0.27577711)


def image_transform(img_u8_hwc, antialias):
    """uint8 [h, w, 3] -> fp32 [3, 224, 224]."""
    x = img_u8_hwc.permute(2, 0, 1).float() / 255.0                        # ToTensor
    _, h, w = x.shape
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = 224, int(224 * long / short)
    rw, rh = (new_short, new_long) if w <= h else (new_long, new_short)
    x = F.interpolate(x[None], size=(rh, rw), mode="bicubic", align_corners=False, antialias=antialias)[0]
    top, left = int(round((rh - 224) / 2.0)), int(round((rw - 224) / 2.0))
    x = x[:, top:top + 224, left:left + 224]
    m, s = torch.tensor(MEAN)[:, None, None], torch.tensor(STD)[:, None, None]
    return (x - m) / s


def video_transform(frames_u8_thwc, flip):
    """uint8 [t, h, w, 3] -> fp32 [3, t, 224, 224]."""
    x = frames_u8_thwc.permute(3, 0, 1, 2).float() / 255.0                 # (C, T, H, W)
    m, s = torch.tensor(MEAN)[:, None, None, None], torch.tensor(STD)[:, None, None, None]
    x = (x - m) / s
    _, _, h, w = x.shape
    if w < h:
        rh, rw = int(math.floor((float(h) / w) * 224)), 224
    else:
        rh, rw = 224, int(math.floor((float(w) / h) * 224))
    x = F.interpolate(x, size=(rh, rw), mode="bilinear", align_corners=False)
    top, left = int(round((rh - 224) / 2.0)), int(round((rw - 224) / 2.0))
    x = x[..., top:top + 224, left:left + 224]
    return x.flip(-1) if flip else x
