"""Cache of kernel-layout (packed, bf16) weights derived from the fp32 reference-layout parameters.

Under ``torch.no_grad()`` a packed tensor is rebuilt only when one of its source parameters changed (tensor
``_version`` / storage pointer, or ``ops.WEIGHT_EPOCH`` -- the fused optimizer rewrites the flat parameter buffer from
a raw-pointer kernel, which no tensor version counter sees), so inference packs once.  With autograd enabled the packing ops (plain torch
elementwise / permute ops on *weights* -- host-side plumbing, not activations) are re-run every forward so that
gradients flow back to the reference-layout parameters.
"""
from __future__ import annotations

from typing import Callable, Dict, Sequence, Tuple

import torch

Tensor = torch.Tensor


class PackCache:
    def __init__(self):
        self._store: Dict[str, Tuple[tuple, object]] = {}

    def get(self, key: str, srcs: Sequence[Tensor], fn: Callable[[], object]):
        if torch.is_grad_enabled() and any(s.requires_grad for s in srcs):
            return fn()
        from . import ops
        sig = (ops.WEIGHT_EPOCH,) + tuple((s.data_ptr(), s._version, s.device) for s in srcs)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = fn()
        self._store[key] = (sig, val)
        return val

    def clear(self):
        self._store.clear()


def bf16c(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).contiguous()


def f32c(t: Tensor) -> Tensor:
    return t.float().contiguous()
