from .attention import FlashAttentionWithRoPE, RoPE2D
from .blocks import ResBlock, RMSNorm, TransVAEBlock
from .conv import ConvFFN
from .upsample import Downsample, Upsample

__all__ = ["FlashAttentionWithRoPE", "RoPE2D", "ResBlock", "RMSNorm", "TransVAEBlock", "ConvFFN", "Downsample",
           "Upsample"]
