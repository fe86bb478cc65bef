"""Downsample / Upsample with DC path -- B200-native mirrors of transvae/modules/upsample.py.

Downsample (upsample.py:55-66): conv3x3 s2 (silu(conv3x3(x))) + conv1x1(pixel_unshuffle(x))  -> 2 launches
Upsample   (upsample.py:116-126): conv3x3(silu(conv3x3(nearest2x(x)))) + pixel_shuffle(conv1x1(x)) -> 2 launches
No pixel (un)shuffle, upsample or add is ever materialised: they are tap tables over phase views.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _taps as T
from .. import kernels as K
from .._pack import bf16c, f32c
from ._base import HotModule


class Downsample(HotModule):
    def __init__(self, in_channels: int, out_channels: int, use_dc_path: bool = True):
        super().__init__()
        from .blocks import _Conv2dParams
        self.use_dc_path = use_dc_path
        self.in_channels, self.out_channels = in_channels, out_channels
        # set by the encoder / decoder when a ResBlock stage follows: the last convolution then also produces the
        # GroupNorm(32) statistics of its output (ops.mtgemm gn_groups)
        self.gn_groups_out = 0
        self.main_path = nn.Sequential(_Conv2dParams(in_channels, in_channels, 3), nn.SiLU(),
                                       _Conv2dParams(in_channels, out_channels, 3))
        if use_dc_path:                      # upsample.py:40-42: no dc_conv parameters without the DC path
            self.dc_conv = _Conv2dParams(in_channels * 4, out_channels, 1)

    def forward_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        B, H, W, C = x.shape
        m0, m2 = self.main_path[0], self.main_path[2]
        dc = self.dc_conv if self.use_dc_path else None
        wdc = dc.weight if dc is not None else None
        bias2 = m2.bias + dc.bias if dc is not None else m2.bias
        if K.needs_grad(x, *self.parameters()):
            from .._autograd import DownsampleFn
            return DownsampleFn.apply(x, T.pack_conv3x3(m0.weight), m0.bias, T.pack_downsample(m2.weight, wdc), bias2)
        w0 = self._packs.get("m0", [m0.weight], lambda: bf16c(T.pack_conv3x3(m0.weight)))
        wd = self._packs.get("down", [m2.weight] + ([wdc] if dc is not None else []),
                             lambda: bf16c(T.pack_downsample(m2.weight, wdc)))
        y = K.mtgemm(T.plan_conv3x3(C), x, w0, out_shape=(B, H, W, C), bias=f32c(m0.bias), act=K.ACT_SILU)
        return K.mtgemm(T.plan_downsample(C, with_dc=dc is not None), y, wd, a1=x if dc is not None else None,
                        out_shape=(B, H // 2, W // 2, self.out_channels), bias=f32c(bias2), gn_groups=self.gn_groups_out)


class Upsample(HotModule):
    def __init__(self, in_channels: int, out_channels: int, use_dc_path: bool = True):
        super().__init__()
        from .blocks import _Conv2dParams
        self.use_dc_path = use_dc_path
        self.in_channels, self.out_channels = in_channels, out_channels
        # set by the encoder / decoder when a ResBlock stage follows: the last convolution then also produces the
        # GroupNorm(32) statistics of its output (ops.mtgemm gn_groups)
        self.gn_groups_out = 0
        self.main_path = nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"),
                                       _Conv2dParams(in_channels, out_channels, 3), nn.SiLU(),
                                       _Conv2dParams(out_channels, out_channels, 3))
        if use_dc_path:                      # upsample.py:101-103
            self.dc_conv = _Conv2dParams(in_channels, out_channels * 4, 1)

    def forward_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        B, H, W, Ci = x.shape
        Co = self.out_channels
        m1, m3 = self.main_path[1], self.main_path[3]
        dc = self.dc_conv if self.use_dc_path else None
        wdc, bdc = (dc.weight, dc.bias) if dc is not None else (None, None)
        if K.needs_grad(x, *self.parameters()):
            from .._autograd import UpsampleFn, pack_upsample_conv1
            return UpsampleFn.apply(x, pack_upsample_conv1(m1.weight), m1.bias, T.pack_upsample_conv2(m3.weight, wdc),
                                    T.bias_upsample_conv2(m3.bias, bdc))
        w1 = self._packs.get("m1", [m1.weight], lambda: bf16c(T.pack_upsample_conv1(m1.weight)))
        b1 = self._packs.get("b1", [m1.bias], lambda: f32c(m1.bias.unsqueeze(0).expand(4, -1)))
        w2 = self._packs.get("m3", [m3.weight] + ([wdc] if dc is not None else []),
                             lambda: bf16c(T.pack_upsample_conv2(m3.weight, wdc)))
        b2 = self._packs.get("b2", [m3.bias] + ([bdc] if dc is not None else []),
                             lambda: f32c(T.bias_upsample_conv2(m3.bias, bdc)))
        y = K.mtgemm(T.plan_upsample_conv1(Ci, Co), x, w1, out_shape=(B, 2 * H, 2 * W, Co), bias=b1, act=K.ACT_SILU)
        return K.mtgemm(T.plan_upsample_conv2(Co, Ci, with_dc=dc is not None), y, w2, a1=x if dc is not None else None,
                        out_shape=(B, 2 * H, 2 * W, Co), bias=b2, gn_groups=self.gn_groups_out)
