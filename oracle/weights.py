"""TEST INFRASTRUCTURE ONLY — deterministic seeded weights.

`seeded_state_dict(shapes, seed)` fills every tensor from a CPU torch.Generator keyed by
(seed, crc32(name)), so the reference (in the build container), the oracle and the CUDA path (on
the GPU box) all see bit-identical fp32 weights without shipping them. It deliberately overwrites
the reference's zero-initialised parameters (SURVEY.md §7 hard part 1), otherwise parity would be
vacuous.
"""
import zlib

import torch


def _is_norm_weight(name, shape):
    if len(shape) != 1:
        return False
    n = name.lower()
    return n.endswith("weight") and any(k in n for k in ("norm", "ln", "layrnorm", ".gn", "bn"))


def seeded_tensor(name, shape, seed, gain=0.8):
    g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    shape = tuple(shape)
    if len(shape) == 0:
        return torch.randn((), generator=g) * 0.5 + 0.5
    if _is_norm_weight(name, shape):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if len(shape) == 1:
        return 0.05 * torch.randn(shape, generator=g)
    n = name.lower()
    if any(k in n for k in ("embed_tokens", "position_embedding", "class_embedding", "temporal_embedding",
                            "query_feat", "query_embed", "level_embed", "gamma_")):
        return 0.5 * torch.randn(shape, generator=g)  # lookup tables / learned tokens, not projections
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return torch.randn(shape, generator=g) * (gain / fan_in ** 0.5)


def seeded_state_dict(shapes, seed, gain=0.8):
    return {k: seeded_tensor(k, v, seed, gain) for k, v in sorted(shapes.items())}


def shapes_of(module_or_sd):
    sd = module_or_sd.state_dict() if hasattr(module_or_sd, "state_dict") else module_or_sd
    return {k: list(v.shape) for k, v in sd.items() if v.dtype.is_floating_point}
