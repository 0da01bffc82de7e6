"""Deterministic synthetic weights / inputs shared by the golden generator and the tests.

Weights are a pure function of (key name, shape, seed): each tensor is drawn from its own
CPU generator seeded by crc32(key) so the result does not depend on module construction
order and can be regenerated on the GPU box (same torch build) instead of being committed.
"""
import zlib

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))
    return g


def synth_tensor(key, shape, seed=0, w_std=0.1, rel_std=0.2, pool_std=0.2, bias_std=0.05, ln_jitter=0.1):
    g = _gen(key, seed)
    r = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    leaf = key.split(".")[-1]
    parent = key.split(".")[-2] if "." in key else ""
    if parent.startswith("norm") and leaf == "weight":
        return 1.0 + ln_jitter * r
    if "rel_pos" in leaf:
        return rel_std * r
    if parent.startswith("pool_"):
        return pool_std * r
    if leaf == "bias":
        return bias_std * r
    if leaf in ("cls_token", "pos_embed_temporal", "object_queries"):
        return 0.5 * r
    return w_std * r


def synth_state(shapes, seed=0, **kw):
    """shapes: dict key -> shape (e.g. {k: v.shape for k, v in module.state_dict().items()})."""
    return {k: synth_tensor(k, s, seed, **kw) for k, s in shapes.items()}


def synth_input(name, shape, seed=0, scale=1.0):
    return scale * torch.randn(tuple(shape), generator=_gen("input:" + name, seed), dtype=torch.float32)
