"""LanguageBind image / video pre-processing on the device.

Mirrors vitron/model/multimodal_encoder/languagebind/image/processing_image.py (`get_image_transform` :15-25,
`LanguageBindImageProcessor.__call__/preprocess` :45-68 -> {"pixel_values": [B, 3, 224, 224]}) and
video/processing_video.py (`get_video_transform` :26-70, frame sampling `np.linspace(0, duration-1, num_frames,
dtype=int)` :91,103 -> {"pixel_values": [B, 3, T, 224, 224]}). Decoding (PIL / decord / opencv) stays on the host;
what moves to the GPU is everything after it: the uint8 frame goes over PCIe once (3 bytes / pixel instead of the
12 bytes / pixel of the fp32 tensor the reference's CPU transform produces) and one fused kernel does
/255 -> resize (short side 224) -> centre crop -> normalise (-> flip), writing the tower's input layout directly.

`antialias`: the reference pins torchvision 0.15.2 (pyproject.toml:16) where Resize on a *tensor* (ToTensor comes
first in the Compose) does not antialias; torchvision >= 0.17 antialiases by default. Both are implemented; the
default follows the pinned version.
"""
import math

import numpy as np
import torch

from . import ops

OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)
MODE_BILINEAR, MODE_BICUBIC, MODE_BICUBIC_AA = 0, 1, 2


def _as_uint8_frames(x, device):
    """PIL image / numpy HWC / uint8 tensor [h,w,3] or [n,h,w,3] -> contiguous uint8 device tensor [n,h,w,3]."""
    if hasattr(x, "convert"):  # PIL
        x = np.asarray(x.convert("RGB"))
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if x.dtype != torch.uint8:
        raise ValueError("device pre-processing takes decoded uint8 frames")
    if x.dim() == 3:
        x = x[None]
    if x.dim() != 4 or x.shape[-1] != 3:
        raise ValueError(f"expected [n, h, w, 3] uint8 frames, got {tuple(x.shape)}")
    return x.to(device, non_blocking=True).contiguous()


def image_resize_geometry(h, w, size=224, crop=224):
    """torchvision Resize(int) (shorter side -> size, other = int(size * long / short)) + CenterCrop offsets."""
    if (w <= h and w == size) or (h <= w and h == size):
        rh, rw = h, w
    elif w < h:
        rw, rh = size, int(size * h / w)
    else:
        rh, rw = size, int(size * w / h)
    return rh, rw, int(round((rh - crop) / 2.0)), int(round((rw - crop) / 2.0))


def video_resize_geometry(h, w, size=224, crop=224):
    """pytorchvideo ShortSideScale (floor) + CenterCropVideo offsets."""
    if w < h:
        rh, rw = int(math.floor((float(h) / w) * size)), size
    else:
        rh, rw = size, int(math.floor((float(w) / h) * size))
    return rh, rw, int(round((rh - crop) / 2.0)), int(round((rw - crop) / 2.0))


def sample_frame_ids(duration, num_frames=8):
    """processing_video.py:91,103."""
    return np.linspace(0, duration - 1, num_frames, dtype=int)


class LanguageBindImageProcessor:
    """Device-side `LanguageBindImageProcessor` (image branch of __call__ / preprocess)."""

    def __init__(self, config=None, tokenizer=None, device="cuda", antialias=False, dtype=torch.float32, **kwargs):
        self.config = config
        self.tokenizer = tokenizer
        self.device = torch.device(device)
        self.antialias = antialias
        self.dtype = dtype
        self.image_mean = OPENAI_DATASET_MEAN
        self.crop_size = {"height": 224, "width": 224}

    def transform_batch(self, frames):
        """uint8 [n, h, w, 3] (equal-sized images) -> [n, 3, 224, 224] with ONE kernel launch."""
        f = _as_uint8_frames(frames, self.device)
        _, h, w, _ = f.shape
        rh, rw, top, left = image_resize_geometry(h, w)
        mode = MODE_BICUBIC_AA if self.antialias else MODE_BICUBIC
        return ops.preprocess_frames(f, rh, rw, top, left, 224, 224, OPENAI_DATASET_MEAN, OPENAI_DATASET_STD, mode,
                                     layout="image", dtype=self.dtype)

    def transform(self, image):
        return self.transform_batch(image)[0]

    def __call__(self, images=None, text=None, context_length=77, return_tensors=None, **kwargs):
        if text is None and images is None:
            raise ValueError("You have to specify either text or images. Both cannot be none.")
        encoding = None
        if text is not None:
            if self.tokenizer is None:
                raise ValueError("text given but the processor has no tokenizer")
            encoding = self.tokenizer(text, max_length=context_length, padding="max_length", truncation=True,
                                      return_tensors=return_tensors, **kwargs)
        if images is not None:
            if torch.is_tensor(images) and images.dim() == 4:   # a batch of equal-sized decoded images: one launch
                feats = self.transform_batch(images)
            else:
                images = images if isinstance(images, list) else [images]
                feats = torch.stack([self.transform(i) for i in images])
            if encoding is not None:
                encoding["pixel_values"] = feats
                return encoding
            return {"pixel_values": feats}
        return encoding

    def preprocess(self, images, return_tensors=None):
        return self.__call__(images=images, return_tensors=return_tensors)


class LanguageBindVideoProcessor:
    """Device-side video transform: decoded uint8 frames [T, H, W, 3] (already sampled with `sample_frame_ids`)
    -> [3, T, 224, 224]. `flip` replaces RandomHorizontalFlipVideo(p=0.5)'s coin (the reference flips at inference
    too, processing_video.py:58): None draws it from `generator`."""

    def __init__(self, config=None, tokenizer=None, device="cuda", dtype=torch.float32, **kwargs):
        self.config = config
        self.tokenizer = tokenizer
        self.device = torch.device(device)
        self.dtype = dtype

    def transform(self, frames, flip=None, generator=None):
        f = _as_uint8_frames(frames, self.device)
        _, h, w, _ = f.shape
        rh, rw, top, left = video_resize_geometry(h, w)
        if flip is None:
            flip = bool(torch.rand(1, generator=generator).item() < 0.5)
        return ops.preprocess_frames(f, rh, rw, top, left, 224, 224, OPENAI_DATASET_MEAN, OPENAI_DATASET_STD, MODE_BILINEAR,
                                     flip=flip, layout="video", dtype=self.dtype)

    def __call__(self, images=None, text=None, context_length=77, return_tensors=None, flip=None, **kwargs):
        if text is None and images is None:
            raise ValueError("You have to specify either text or images. Both cannot be none.")
        encoding = None
        if text is not None:
            if self.tokenizer is None:
                raise ValueError("text given but the processor has no tokenizer")
            encoding = self.tokenizer(text, max_length=context_length, padding="max_length", truncation=True,
                                      return_tensors=return_tensors, **kwargs)
        if images is not None:
            images = images if isinstance(images, list) else [images]
            feats = torch.stack([self.transform(v, flip=flip) for v in images])
            if encoding is not None:
                encoding["pixel_values"] = feats
                return encoding
            return {"pixel_values": feats}
        return encoding

    def preprocess(self, images, return_tensors=None, flip=None):
        return self.__call__(images=images, return_tensors=return_tensors, flip=flip)
