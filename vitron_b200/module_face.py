"""The slice of the `torch.nn.Module` surface that the reference's callers touch on the drop-in classes
(builder.py:152-163 `tower.to(device=..., dtype=...)`, inference_image.py:27 `model.device`, `.eval()`, `next(model.parameters())`,
`state_dict()` for checkpoint round trips). The drop-ins own packed bf16 device tensors rather than nn.Parameters, so `.to()` can
only confirm the placement they were built with: asking for another device is an error, never a silent copy or a CPU path."""
import torch


def check_to(device, args, kwargs):
    want_dev = kwargs.get("device")
    want_dtype = kwargs.get("dtype")
    for a in args:
        if isinstance(a, (str, torch.device)):
            want_dev = a
        elif isinstance(a, torch.dtype):
            want_dtype = a
    if want_dev is not None:
        wd = torch.device(want_dev)
        same = wd.type == device.type and (wd.index is None or device.index is None or wd.index == device.index)
        if not same:
            raise ValueError(f"vitron_b200 modules live on the device they were built on ({device}); rebuild for {wd}")
    if want_dtype is not None and want_dtype not in (torch.float16, torch.bfloat16):
        raise ValueError(f"vitron_b200 computes in bf16 (requested {want_dtype}); fp16 requests from the reference's loaders are accepted "
                         "and served in bf16")


class ModuleFace:
    """Mixin: expects `self.device`; subclasses provide `parameters()` and `state_dict()`."""
    dtype = torch.bfloat16
    training = False

    def to(self, *args, **kwargs):
        check_to(self.device, args, kwargs)
        return self

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def half(self):
        return self

    def bfloat16(self):
        return self

    def eval(self):
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("inference path only (SURVEY.md §2: training is out of scope)")
        return self

    def requires_grad_(self, flag=False):
        return self

    def named_parameters(self):
        return iter(self.state_dict().items())
