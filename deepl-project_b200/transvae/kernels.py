"""Autograd-aware kernel layer used by the modules.

Without autograd (inference) each function is a direct launch from ``ops``.  With autograd enabled the same
entry points route through ``torch.autograd.Function`` wrappers (``_autograd.py``) whose backward passes are
again hand-written kernels.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops
from ._lib import ACT_GELU, ACT_NONE, ACT_SILU  # noqa: F401

Tensor = torch.Tensor


def _needs_grad(*ts) -> bool:
    for t in ts:
        if isinstance(t, torch.Tensor) and not t.is_cuda:
            raise RuntimeError("TransVAE B200 path needs CUDA tensors on an sm_100 device (there is no CPU fallback)")
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


def _ag():
    from . import _autograd
    return _autograd


def nchw_to_nhwc(x: Tensor, cpad: int) -> Tensor:
    if _needs_grad(x):
        return _ag().NchwToNhwc.apply(x, cpad)
    return ops.nchw_to_nhwc(x, cpad)


def nhwc_to_nchw(x: Tensor, c: Optional[int] = None) -> Tensor:
    if _needs_grad(x):
        return _ag().NhwcToNchw.apply(x, c)
    return ops.nhwc_to_nchw(x, c)


def needs_grad(*ts) -> bool:
    """True when the training path (block-level autograd Functions) must be taken."""
    return _needs_grad(*ts)


def conv_in(x: Tensor, w: Tensor, b: Optional[Tensor], gn_groups: int = 0) -> Tensor:
    if _needs_grad(x, w, b):
        return _ag().ConvInFn.apply(x, w, b)
    return ops.conv_in(x, w, b, gn_groups=gn_groups)


def mtgemm(plan, a0: Tensor, w: Tensor, **kw) -> Tensor:
    if _needs_grad(a0, w, kw.get("a1"), kw.get("bias"), kw.get("residual")):
        raise RuntimeError("internal: a bare mtgemm launch was reached with autograd enabled; the training path goes "
                           "through the block-level Functions of transvae._autograd")
    return ops.mtgemm(plan, a0, w, **kw)


def linear(x: Tensor, w: Tensor, plan, **kw) -> Tensor:
    lead = x.shape[:-1]
    m = x.numel() // x.shape[-1]
    res = kw.pop("residual", None)
    if res is not None:
        res = res.reshape(1, 1, m, w.shape[0])
    y = mtgemm(plan, x.reshape(1, 1, m, x.shape[-1]), w, out_shape=(1, 1, m, w.shape[0]), residual=res, **kw)
    return y.reshape(*lead, w.shape[0])


def groupnorm_silu(x: Tensor, gamma: Tensor, beta: Tensor, silu: bool = True) -> Tensor:
    if _needs_grad(x, gamma, beta):
        return _ag().GroupNormSilu.apply(x, gamma, beta, silu, gn_sums_of(x))
    # statistics left on the tensor by the convolution that produced it (ops.mtgemm(..., gn_groups=32)): apply pass only
    return ops.groupnorm_silu(x, gamma, beta, silu=silu, sums=gn_sums_of(x))


def gn_sums_of(x: Tensor, groups: int = 32) -> Optional[Tensor]:
    s = getattr(x, "_gn_sums", None)
    if s is not None and tuple(s.shape) == (x.shape[0], groups, 2) and s.device == x.device and s.dtype == torch.float64:
        return s
    return None


def dwconv3x3(u: Tensor, w9c: Tensor, bias: Optional[Tensor]) -> Tensor:
    if _needs_grad(u, w9c, bias):
        raise RuntimeError("internal: the depthwise ConvFFN trains through _autograd.FfnDwFn")
    return ops.dwconv3x3(u, w9c, bias, flip=False, add_input=True)


def rms_norm(x: Tensor, w: Tensor) -> Tensor:
    """RMSNorm over the last axis of a contiguous bf16 tensor (one kernel; differentiable)."""
    if _needs_grad(x, w):
        return _ag().RmsNormFn.apply(x, w)
    return ops.token_norm_fwd(x, w, 0)


def row_stats(x: Tensor, w1: Optional[Tensor] = None, mode: Optional[int] = None):
    return ops.row_stats(x, w1, mode)


def attention(qkv: Tensor, B: int, S: int, C: int) -> Tensor:
    return ops.attn_fwd(qkv, B, S, C)[0]


def reparam(mu: Tensor, logvar: Tensor, eps: Tensor, patched: bool) -> Tuple[Tensor, Tensor, Tensor]:
    if _needs_grad(mu, logvar):
        return _ag().Reparam.apply(mu, logvar, eps, patched)
    return ops.reparam(mu, logvar, eps, patched)


def loss_sums(recon: Tensor, target: Tensor, mu: Tensor, logvar: Tensor, patched: bool, clip=(-30.0, 20.0)):
    """(sum |f(recon) - target|, sum KL terms, #non-finite terms) as 0-dim fp32 tensors."""
    if _needs_grad(recon, mu, logvar):
        return _ag().LossSums.apply(recon, target, mu, logvar, patched, float(clip[0]), float(clip[1]))
    acc = ops.loss_sums(recon, target, mu, logvar, patched, clip)
    return acc[0], acc[1], acc[2]
