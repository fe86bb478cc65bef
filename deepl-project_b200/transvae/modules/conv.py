"""ConvFFN -- B200-native mirror of transvae/modules/conv.py: conv_type='full' (the variant the shipped configs
reach; conv.py:30, blocks.py:119-123) and conv_type='depthwise' (conv.py:42-50, reachable by constructing the module
directly: u + depthwise3x3(u) as one HBM-bound stencil kernel between the two projections).

forward (conv.py:79-105):  u = gelu(proj_in(x));  u = u + conv(u),  conv = 1x1 -> GELU -> 3x3 -> GELU -> 1x1;
out = proj_out(u).  Five tcgen05 GEMM launches with bias / GELU / residual fused into the epilogues; the
preceding RMSNorm is folded into proj_in (row scale 1/rms in the epilogue, weight into the columns).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _taps as T
from .. import kernels as K
from .._pack import bf16c, f32c
from ._base import HotModule


class ConvFFN(HotModule):
    def __init__(self, dim: int, mlp_ratio: float = 1.0, conv_type: str = "full", dropout: float = 0.0):
        super().__init__()
        from .blocks import _Conv2dParams, _LinearParams
        if conv_type not in ("full", "depthwise"):
            raise ValueError(f"Unknown conv_type: {conv_type}")            # conv.py:62
        self.dim, self.conv_type = dim, conv_type
        hidden = int(dim * mlp_ratio * 4)
        mid = int(dim * mlp_ratio)
        self.hidden_dim, self.conv_hidden = hidden, mid
        self.proj_in = _LinearParams(dim, hidden)
        if conv_type == "depthwise":
            # nn.Conv2d(hidden, hidden, 3, padding=1, groups=hidden): weight [hidden, 1, 3, 3] (conv.py:44-50)
            self.conv = _Conv2dParams(hidden, hidden, 3, groups=hidden)
        else:
            self.conv = nn.Sequential(_Conv2dParams(hidden, mid, 1), nn.GELU(), _Conv2dParams(mid, mid, 3), nn.GELU(),
                                      _Conv2dParams(mid, hidden, 1))
        self.proj_out = _LinearParams(hidden, dim)
        self.dropout = nn.Dropout(dropout)

    def forward_fused(self, x: torch.Tensor, w2, add_residual: bool = True) -> torch.Tensor:
        """x + ffn(RMSNorm(x; w2)) for NHWC bf16 x; ``w2 is None`` = no norm (bare module)."""
        B, H, W, C = x.shape
        M = B * H * W
        hid, mid = self.hidden_dim, self.conv_hidden
        if self.conv_type == "depthwise":
            return self._forward_depthwise(x, w2, add_residual)
        if K.needs_grad(x, w2, *self.parameters()):
            if w2 is None or not add_residual:
                raise NotImplementedError("bare ConvFFN (no RMSNorm / no residual) is an inference-only hook")
            from .._autograd import FfnFn
            c0, c2, c4 = self.conv[0], self.conv[2], self.conv[4]
            return FfnFn.apply(x, w2, self.proj_in.weight, self.proj_in.bias, c0.weight, c0.bias,
                               c2.weight, c2.bias, c4.weight, c4.bias,
                               self.proj_out.weight, self.proj_out.bias)
        xf = x.reshape(M, C)
        if w2 is not None:
            w_in = self._packs.get("in", [self.proj_in.weight, w2], lambda: bf16c(self.proj_in.weight * w2))
            rstd, _ = K.row_stats(x)
        else:
            w_in = self._packs.get("in_raw", [self.proj_in.weight], lambda: bf16c(self.proj_in.weight))
            rstd = None
        u = K.linear(xf, w_in, T.plan_linear(C), bias=f32c(self.proj_in.bias), row_scale=rstd, act=K.ACT_GELU)
        c0, c2, c4 = self.conv[0], self.conv[2], self.conv[4]
        w0 = self._packs.get("c0", [c0.weight], lambda: bf16c(T.pack_conv1x1(c0.weight)))
        w2_ = self._packs.get("c2", [c2.weight], lambda: bf16c(T.pack_conv3x3(c2.weight)))
        w4 = self._packs.get("c4", [c4.weight], lambda: bf16c(T.pack_conv1x1(c4.weight)))
        t = K.linear(u, w0, T.plan_linear(hid), bias=f32c(c0.bias), act=K.ACT_GELU)
        t = K.mtgemm(T.plan_conv3x3(mid), t.reshape(B, H, W, mid), w2_, out_shape=(B, H, W, mid), bias=f32c(c2.bias),
                     act=K.ACT_GELU)
        u = K.linear(t.reshape(M, mid), w4, T.plan_linear(mid), bias=f32c(c4.bias), residual=u)
        wo = self._packs.get("out", [self.proj_out.weight], lambda: bf16c(self.proj_out.weight))
        y = K.linear(u, wo, T.plan_linear(hid), bias=f32c(self.proj_out.bias), residual=xf if add_residual else None)
        return y.reshape(B, H, W, C)

    def _dw_weight(self) -> torch.Tensor:
        """[hidden, 1, 3, 3] -> fp32 [9, hidden] (tap-major) for ``tvae_dwconv3x3`` (a differentiable torch re-layout)."""
        return self.conv.weight.reshape(self.hidden_dim, 9).t()

    def _forward_depthwise(self, x: torch.Tensor, w2, add_residual: bool) -> torch.Tensor:
        B, H, W, C = x.shape
        M, hid = B * H * W, self.hidden_dim
        if K.needs_grad(x, w2, *self.parameters()):
            if w2 is None or not add_residual:
                raise NotImplementedError("bare ConvFFN (no RMSNorm / no residual) is an inference-only hook")
            from .._autograd import FfnDwFn
            return FfnDwFn.apply(x, w2, self.proj_in.weight, self.proj_in.bias, self._dw_weight().contiguous(), self.conv.bias,
                                 self.proj_out.weight, self.proj_out.bias)
        xf = x.reshape(M, C)
        if w2 is not None:
            w_in = self._packs.get("in", [self.proj_in.weight, w2], lambda: bf16c(self.proj_in.weight * w2))
            rstd, _ = K.row_stats(x)
        else:
            w_in = self._packs.get("in_raw", [self.proj_in.weight], lambda: bf16c(self.proj_in.weight))
            rstd = None
        u = K.linear(xf, w_in, T.plan_linear(C), bias=f32c(self.proj_in.bias), row_scale=rstd, act=K.ACT_GELU)
        wdw = self._packs.get("dw", [self.conv.weight], lambda: f32c(self._dw_weight()))
        u = K.dwconv3x3(u.reshape(B, H, W, hid), wdw, f32c(self.conv.bias))           # u + dwconv(u) + bias
        wo = self._packs.get("out", [self.proj_out.weight], lambda: bf16c(self.proj_out.weight))
        y = K.linear(u.reshape(M, hid), wo, T.plan_linear(hid), bias=f32c(self.proj_out.bias),
                     residual=xf if add_residual else None)
        return y.reshape(B, H, W, C)

    def forward_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward_fused(x, None, add_residual=False)
