"""FlashAttentionWithRoPE / RoPE2D -- B200-native mirrors of transvae/modules/attention.py.

forward (attention.py:55-104):  q,k,v = to_{q,k,v}(LayerNorm_{q,k,v}(x));  RoPE(q), RoPE(k);
SDPA(scale = head_dim^-0.5);  proj.  Here: one statistics pass, ONE [3C, C] GEMM whose epilogue applies the folded
norms, the reference's (non-orthogonal) RoPE and the softmax scale, the tcgen05 flash kernel, and the output
projection with the residual add fused.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import _taps as T
from .. import kernels as K
from .._pack import bf16c, f32c
from ._base import HotModule


class RoPE2D(nn.Module):
    """Holds the ``inv_freq`` buffer (attention.py:126-130).  The rotation itself runs in the QKV GEMM epilogue."""

    def __init__(self, dim: int, max_resolution: int = 4096):
        super().__init__()
        assert dim % 2 == 0, "Dimension must be even for RoPE"
        self.dim, self.max_resolution = dim, max_resolution
        dpa = dim // 2
        self.register_buffer("inv_freq", 1.0 / (10000 ** (torch.arange(0, dpa, 2).float() / dpa)))

    def table(self, H: int, W: int) -> torch.Tensor:
        return T.rope_table(H, W, self.inv_freq)


class FlashAttentionWithRoPE(HotModule):
    def __init__(self, dim: int, head_dim: int = 64, use_rope: bool = True, dropout: float = 0.0):
        super().__init__()
        from .blocks import _LinearParams, _NormParams
        if head_dim != 64:
            raise NotImplementedError("the sm_100a attention kernel is specialised for head_dim 64 (every reference config)")
        self.dim, self.head_dim, self.num_heads = dim, head_dim, dim // head_dim
        self.scale = head_dim ** -0.5
        self.use_rope = use_rope
        self.norm_q, self.norm_k, self.norm_v = _NormParams(dim), _NormParams(dim), _NormParams(dim)
        self.to_q = _LinearParams(dim, dim, bias=False)
        self.to_k = _LinearParams(dim, dim, bias=False)
        self.to_v = _LinearParams(dim, dim, bias=False)
        self.proj = _LinearParams(dim, dim)
        self.dropout = nn.Dropout(dropout)
        if use_rope:                          # attention.py:50-53: no RoPE2D submodule (and no inv_freq buffer) otherwise
            self.rope = RoPE2D(head_dim)
        self._rope_cache = {}

    def _rope_tab(self, H: int, W: int) -> torch.Tensor:
        if not self.use_rope:
            # identity rotation (cos 1, sin 0): the QKV epilogue and the attention backward then leave q, k untouched
            dev = self.to_q.weight.device
            key = (H, W, dev, -1)
            tab = self._rope_cache.get(key)
            if tab is None:
                tab = torch.zeros(max(H, W), self.head_dim // 4, 2, dtype=torch.float32, device=dev)
                tab[..., 0] = 1.0
                self._rope_cache = {key: tab}
            return tab
        key = (H, W, self.rope.inv_freq.device, self.rope.inv_freq._version)
        tab = self._rope_cache.get(key)
        if tab is None:
            tab = self.rope.table(H, W)
            self._rope_cache = {key: tab}
        return tab

    def _folded(self, w1: torch.Tensor):
        srcs = [self.to_q.weight, self.to_k.weight, self.to_v.weight, self.norm_q.weight, self.norm_q.bias,
                self.norm_k.weight, self.norm_k.bias, self.norm_v.weight, self.norm_v.bias, w1]

        def make():
            W, cs, b = T.fold_qkv(self.to_q.weight, self.to_k.weight, self.to_v.weight, self.norm_q.weight,
                                  self.norm_q.bias, self.norm_k.weight, self.norm_k.bias, self.norm_v.weight,
                                  self.norm_v.bias, w1)
            return bf16c(W), f32c(cs), f32c(b)
        return self._packs.get("qkv", srcs, make)

    def forward_fused(self, x: torch.Tensor, w1: torch.Tensor, add_residual: bool = True, rms: bool = True) -> torch.Tensor:
        """x + proj(SDPA(...)) with the preceding RMSNorm (weight ``w1``) folded in.  x: NHWC bf16.
        ``rms=False``: no RMSNorm in front (bare module); ``add_residual=False``: return the branch only."""
        B, H, W, C = x.shape
        S = H * W
        if K.needs_grad(x, w1, *self.parameters()):
            if not (add_residual and rms):
                raise NotImplementedError("bare attention (no RMSNorm / no residual) is an inference-only hook")
            from .._autograd import AttnFn, fold_qkv_affine
            wg, bg = fold_qkv_affine(self.to_q.weight, self.to_k.weight, self.to_v.weight, self.norm_q.weight,
                                       self.norm_q.bias, self.norm_k.weight, self.norm_k.bias, self.norm_v.weight,
                                       self.norm_v.bias)
            return AttnFn.apply(x, w1, wg, bg, self.proj.weight, self.proj.bias, self._rope_tab(H, W), self.scale)
        wqkv, colsum, bias = self._folded(w1)
        a, b = K.row_stats(x, w1.detach(), 1 if rms else 2)
        rope = (self._rope_tab(H, W), C, H, W, self.scale * math.log2(math.e))
        qkv = K.linear(x.reshape(B * S, C), wqkv, T.plan_linear(C), bias=bias, row_scale=a, row_shift=b,
                       col_sum=colsum, rope=rope)
        o = K.attention(qkv, B, S, C)
        wp = self._packs.get("proj", [self.proj.weight], lambda: bf16c(self.proj.weight))
        y = K.linear(o.reshape(B * S, C), wp, T.plan_linear(C), bias=f32c(self.proj.bias),
                     residual=x.reshape(B * S, C) if add_residual else None)
        return y.reshape(B, H, W, C)

    def forward_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        """Reference semantics of the bare module (no RMSNorm before, no residual after): attn(x)."""
        ones = torch.ones(self.dim, device=x.device)
        return self.forward_fused(x, ones, add_residual=False, rms=False)
