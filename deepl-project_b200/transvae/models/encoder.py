"""TransVAEEncoder -- B200-native mirror of transvae/models/encoder.py.

conv_in -> [ResBlock x d0] -> Down -> [ResBlock x d1] -> Down -> [TransVAEBlock x d2..] with a Downsample after
every stage but the last (encoder.py:101-126; the number of stride-2 stages is len(depths)-1, the
``compression_ratio`` argument is ignored exactly like in the reference).
Input: NCHW float image.  Output of ``forward``: NCHW float features (reference signature); the VAE calls
``forward_features`` which keeps the NHWC bf16 activation for the fused mu/logvar heads.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from .. import kernels as K
from ..modules.blocks import ResBlock, TransVAEBlock, _Conv2dParams
from ..modules.upsample import Downsample


def _run_block(block, h, ckpt: bool):
    if ckpt and torch.is_grad_enabled():
        return torch.utils.checkpoint.checkpoint(block.forward_nhwc, h, use_reentrant=False)
    return block.forward_nhwc(h)


class TransVAEEncoder(nn.Module):
    def __init__(self, input_channels: int = 3, latent_dim: int = 32, depths: List[int] = [3, 3, 3, 4, 6],
                 base_dims: List[int] = [192, 192, 384, 768, 1536], compression_ratio: int = 16,
                 mlp_ratio: float = 1.0, head_dim: int = 64, use_rope: bool = True, use_conv_ffn: bool = True,
                 use_dc_path: bool = True):
        super().__init__()
        self.num_stages = len(depths)
        self.depths, self.base_dims, self.compression_ratio = list(depths), list(base_dims), compression_ratio
        self.conv_in = _Conv2dParams(input_channels, base_dims[0], 3)
        self.stages = nn.ModuleList()
        self.downsamples = nn.ModuleList()
        for i in range(self.num_stages):
            d = base_dims[i]
            if i < 2:
                blocks = nn.ModuleList([ResBlock(d, d) for _ in range(depths[i])])
            else:
                blocks = nn.ModuleList([TransVAEBlock(dim=d, mlp_ratio=mlp_ratio, head_dim=head_dim, use_rope=use_rope,
                                                      use_conv_ffn=use_conv_ffn) for _ in range(depths[i])])
            self.stages.append(blocks)
            if i < self.num_stages - 1:
                self.downsamples.append(Downsample(d, base_dims[i + 1], use_dc_path=use_dc_path))
                if i + 1 < 2:                 # the next stage is a ResBlock stage: hand it the statistics of its input
                    self.downsamples[-1].gn_groups_out = 32
        self.gradient_checkpointing = False

    def enable_gradient_checkpointing(self):
        self.gradient_checkpointing = True

    def forward_features(self, x: torch.Tensor, trace: dict = None) -> torch.Tensor:
        """NCHW float image -> NHWC bf16 features [B, H/f, W/f, C_last]."""
        h = K.conv_in(x, self.conv_in.weight, self.conv_in.bias, gn_groups=32)
        if trace is not None:
            trace["encoder.conv_in"] = h
        ckpt = self.gradient_checkpointing and self.training
        for i, stage in enumerate(self.stages):
            for j, block in enumerate(stage):
                h = _run_block(block, h, ckpt)
                if trace is not None:
                    trace[f"encoder.stages.{i}.{j}"] = h
            if i < len(self.downsamples):
                h = self.downsamples[i].forward_nhwc(h)
                if trace is not None:
                    trace[f"encoder.downsamples.{i}"] = h
        return h

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return K.nhwc_to_nchw(self.forward_features(x)).to(x.dtype if x.is_floating_point() else torch.float32)
