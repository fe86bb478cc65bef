"""TransVAEDecoder -- B200-native mirror of transvae/models/decoder.py.

conv_in -> TransVAEBlock stages -> ... -> ResBlock stages with an Upsample after every stage but the last, then
GroupNorm(32) -> SiLU -> conv_out (decoder.py:102-132).  Input: NCHW float latent z; output: NCHW fp32
reconstruction logits (unbounded, as in the reference).
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from .. import _taps as T
from .. import kernels as K
from .._pack import PackCache, bf16c, f32c
from ..modules.blocks import ResBlock, TransVAEBlock, _Conv2dParams, _NormParams
from ..modules.upsample import Upsample
from .encoder import _run_block


class TransVAEDecoder(nn.Module):
    def __init__(self, latent_dim: int = 32, output_channels: int = 3, depths: List[int] = [6, 4, 3, 3, 3],
                 base_dims: List[int] = [1536, 768, 384, 192, 192], compression_ratio: int = 16,
                 mlp_ratio: float = 1.0, head_dim: int = 64, use_rope: bool = True, use_conv_ffn: bool = True,
                 use_dc_path: bool = True):
        super().__init__()
        self.num_stages = len(depths)
        self.depths, self.base_dims = list(depths), list(base_dims)
        self.latent_dim, self.output_channels = latent_dim, output_channels
        self.conv_in = _Conv2dParams(latent_dim, base_dims[0], 3)
        self.stages = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        n_tr = self.num_stages - 2
        for i in range(self.num_stages):
            d = base_dims[i]
            if i < n_tr:
                blocks = nn.ModuleList([TransVAEBlock(dim=d, mlp_ratio=mlp_ratio, head_dim=head_dim, use_rope=use_rope,
                                                      use_conv_ffn=use_conv_ffn) for _ in range(depths[i])])
            else:
                blocks = nn.ModuleList([ResBlock(d, d) for _ in range(depths[i])])
            self.stages.append(blocks)
            if i < self.num_stages - 1:
                self.upsamples.append(Upsample(d, base_dims[i + 1], use_dc_path=use_dc_path))
                if i + 1 >= n_tr:             # a ResBlock stage follows: hand it the statistics of its input
                    self.upsamples[-1].gn_groups_out = 32
        self.norm_out = _NormParams(base_dims[-1], 32)
        self.conv_out = _Conv2dParams(base_dims[-1], output_channels, 3)
        self.gradient_checkpointing = False
        object.__setattr__(self, "_packs", PackCache())

    def _apply(self, fn, *a, **kw):
        self._packs.clear()
        return super()._apply(fn, *a, **kw)

    def enable_gradient_checkpointing(self):
        self.gradient_checkpointing = True

    def forward(self, z: torch.Tensor, trace: dict = None) -> torch.Tensor:
        B, D, H, W = z.shape
        cpad = (D + 63) // 64 * 64
        C0 = self.base_dims[0]
        train = K.needs_grad(z, *self.parameters())
        zn = K.nchw_to_nhwc(z, cpad)
        if train:
            from .._autograd import Conv3x3Fn
            h = Conv3x3Fn.apply(zn, T.pack_conv3x3(self.conv_in.weight, cin_pad=cpad), self.conv_in.bias)
        else:
            w_in = self._packs.get("in", [self.conv_in.weight], lambda: bf16c(T.pack_conv3x3(self.conv_in.weight, cin_pad=cpad)))
            h = K.mtgemm(T.plan_conv3x3(cpad), zn, w_in, out_shape=(B, H, W, C0), bias=f32c(self.conv_in.bias))
        if trace is not None:
            trace["decoder.conv_in"] = h
        ckpt = self.gradient_checkpointing and self.training
        for i, stage in enumerate(self.stages):
            for j, block in enumerate(stage):
                h = _run_block(block, h, ckpt)
                if trace is not None:
                    trace[f"decoder.stages.{i}.{j}"] = h
            if i < len(self.upsamples):
                h = self.upsamples[i].forward_nhwc(h)
                if trace is not None:
                    trace[f"decoder.upsamples.{i}"] = h
        h = K.groupnorm_silu(h, self.norm_out.weight, self.norm_out.bias)
        Bh, Hh, Wh, Ch = h.shape
        oc = self.output_channels
        npad = (oc + 63) // 64 * 64
        if train:
            from .._autograd import HeadFn
            return HeadFn.apply(h, T.pack_conv3x3(self.conv_out.weight, cout_pad=npad),
                                torch.nn.functional.pad(self.conv_out.bias, (0, npad - oc)), oc)
        w_out = self._packs.get("out", [self.conv_out.weight], lambda: bf16c(T.pack_conv3x3(self.conv_out.weight, cout_pad=npad)))
        b_out = self._packs.get("bout", [self.conv_out.bias],
                                lambda: f32c(torch.nn.functional.pad(self.conv_out.bias, (0, npad - oc))))
        return K.mtgemm(T.plan_conv3x3(Ch), h, w_out, bias=b_out, out_f32_shape=(Bh, oc, Hh, Wh), out_n=oc)
