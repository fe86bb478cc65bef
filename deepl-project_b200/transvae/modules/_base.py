"""Shared plumbing for the hot-path modules.

Internal activations are NHWC bf16 tensors [B, H, W, C].  Every module exposes

* ``forward_nhwc(x)``: the hot path (NHWC bf16 in / out), used by the encoder / decoder, and
* ``forward(x)``: the reference's public signature (NCHW float in, NCHW float out), which converts at the
  boundary with the layout kernels and calls ``forward_nhwc``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import kernels as K
from .._pack import PackCache


class HotModule(nn.Module):
    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_packs", PackCache())

    def forward_nhwc(self, x: torch.Tensor) -> torch.Tensor:  # pragma: no cover
        raise NotImplementedError

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Reference signature: x [B, C, H, W] float -> [B, C', H', W'] in x.dtype."""
        y = self.forward_nhwc(K.nchw_to_nhwc(x, x.shape[1]))
        return K.nhwc_to_nchw(y).to(x.dtype)

    def _apply(self, fn, *a, **kw):
        self._packs.clear()
        return super()._apply(fn, *a, **kw)
