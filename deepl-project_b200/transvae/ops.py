"""Kernel launch wrappers: torch tensors in, torch tensors out, all compute in libtransvae_sm100.so.

PyTorch is used for device memory (caching allocator) and the current CUDA stream only.  Every function
raises if its inputs are not CUDA tensors -- there is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_SILU, MtGemmDesc  # noqa: F401
from ._taps import Plan

Tensor = torch.Tensor
BF16 = torch.bfloat16

# launch counter (bench.py reports it as gpu_launches)
LAUNCHES = 0
# optional per-launch device timing of the tensor-core kernel (bench.py roofline): list of
# (name, algorithmic_flops, start_event, end_event); None = off
PROFILE = None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("TransVAE B200 ops need CUDA tensors (no CPU fallback)")


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _set_view(v, t: Optional[Tensor], split: bool) -> None:
    if t is None:
        v.ptr = None
        return
    assert t.dtype == BF16 and t.is_contiguous() and t.dim() == 4, (t.dtype, t.shape, t.is_contiguous())
    v.ptr = t.data_ptr()
    v.B, v.H, v.W, v.C = t.shape
    v.split = 1 if split else 0


def mtgemm(plan: Plan, a0: Tensor, w: Tensor, *, a1: Optional[Tensor] = None, out: Optional[Tensor] = None,
           out_shape: Optional[Sequence[int]] = None, bias: Optional[Tensor] = None, act: int = ACT_NONE,
           residual: Optional[Tensor] = None, row_scale: Optional[Tensor] = None, row_shift: Optional[Tensor] = None,
           col_sum: Optional[Tensor] = None, rope: Optional[Tuple[Tensor, int, int, int, float]] = None,
           out_f32: Optional[Tensor] = None, out_f32_shape: Optional[Sequence[int]] = None, out_n: int = 0) -> Tensor:
    """Launch ``tvae_mtgemm``.  a0 / a1 / out / residual are NHWC bf16 4-D tensors (flat matrices as
    [1, 1, M, K]); ``w`` is the packed bf16 [N, K_total] weight; ``bias`` fp32 [phases, N] (or [N])."""
    _need_cuda(a0, w, a1, out, bias, residual, row_scale, row_shift, col_sum, out_f32)
    assert w.dtype == BF16 and w.is_contiguous() and w.shape[1] == plan.k_total, (w.shape, plan.k_total)
    d = MtGemmDesc()
    _set_view(d.a0, a0, plan.a0_split)
    _set_view(d.a1, a1, plan.a1_split)
    n_total = w.shape[0]
    if out_f32 is None and out_f32_shape is not None:
        out_f32 = torch.empty(tuple(out_f32_shape), dtype=torch.float32, device=a0.device)
    if out_f32 is None:
        if out is None:
            out = torch.empty(tuple(out_shape), dtype=BF16, device=a0.device)
        _set_view(d.out, out, plan.out_split)
    else:
        assert out_f32.dtype == torch.float32 and out_f32.is_contiguous()
        d.out.ptr = None
    _set_view(d.res, residual, plan.out_split)
    if residual is not None:
        assert residual.shape == out.shape
    d.w = w.data_ptr()
    d.n_total, d.k_total = n_total, plan.k_total
    d.num_phases = plan.num_phases
    for ph, taps in enumerate(plan.phases):
        d.ntaps[ph] = len(taps)
        d.out_p[ph] = plan.out_p[ph]
        d.out_c_off[ph] = plan.out_c_off[ph]
        for i, t in enumerate(taps):
            dt = d.taps[ph][i]
            dt.map, dt.c_off, dt.dw, dt.p, dt.dh, dt.kblocks, dt.wk_off = t.map, t.c_off, t.dw, t.p, t.dh, t.kblocks, t.wk_off
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == plan.num_phases * n_total, \
            (bias.shape, plan.num_phases, n_total)
    d.bias = _ptr(bias)
    d.act = act
    for name, t in (("row_scale", row_scale), ("row_shift", row_shift), ("col_sum", col_sum)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous()
        setattr(d, name, _ptr(t))
    if rope is not None:
        tab, rc, rh, rw, qs = rope
        assert tab.dtype == torch.float32 and tab.is_contiguous() and tab.shape[0] >= max(rh, rw)
        d.rope_tab, d.rope_C, d.rope_H, d.rope_W, d.q_scale = tab.data_ptr(), rc, rh, rw, qs
    else:
        d.rope_tab, d.q_scale = None, 1.0
    d.out_f32 = _ptr(out_f32)
    d.out_n = out_n
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(_lib.load().tvae_mtgemm(C.byref(d), _stream()), f"tvae_mtgemm[{plan.name}]")
    if PROFILE is not None:
        e1.record()
        o = out if out_f32 is None else out_f32
        n_real = out_n if out_f32 is not None else n_total
        m_out = o.numel() // (o.shape[1] if out_f32 is not None else o.shape[-1])
        PROFILE.append((plan.name, 2.0 * m_out * n_real * plan.algo_k, e0, e1))
    _count()
    return out if out_f32 is None else out_f32


def linear(x: Tensor, w: Tensor, plan: Plan, **kw) -> Tensor:
    """Flat token GEMM: x [..., K] bf16 -> [..., N]."""
    lead = x.shape[:-1]
    m = x.numel() // x.shape[-1]
    res = kw.pop("residual", None)
    if res is not None:
        res = res.reshape(1, 1, m, w.shape[0])
    y = mtgemm(plan, x.reshape(1, 1, m, x.shape[-1]), w, out_shape=(1, 1, m, w.shape[0]), residual=res, **kw)
    return y.reshape(*lead, w.shape[0])


def attn_fwd(qkv: Tensor, B: int, S: int, C_: int, need_lse: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
    _need_cuda(qkv)
    assert qkv.dtype == BF16 and qkv.is_contiguous() and qkv.numel() == B * S * 3 * C_
    out = torch.empty(B, S, C_, dtype=BF16, device=qkv.device)
    lse = torch.empty(B, C_ // 64, S, dtype=torch.float32, device=qkv.device) if need_lse else None
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(_lib.load().tvae_attn_fwd(qkv.data_ptr(), out.data_ptr(), _ptr(lse), B, S, C_, _stream()), "tvae_attn_fwd")
    if PROFILE is not None:
        e1.record()
        PROFILE.append(("attn_fwd", 4.0 * B * S * S * C_, e0, e1))
    _count()
    return out, lse


def conv_in(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """NCHW fp32 image -> NHWC bf16 features."""
    _need_cuda(x, w, b)
    x = x.float().contiguous()
    B, Cin, H, W = x.shape
    out = torch.empty(B, H, W, w.shape[0], dtype=BF16, device=x.device)
    wf = w.float().contiguous()
    bf = None if b is None else b.float().contiguous()
    _lib.check(_lib.load().tvae_conv_in(x.data_ptr(), wf.data_ptr(), _ptr(bf), out.data_ptr(), B, Cin, H, W, w.shape[0],
                                        _stream()), "tvae_conv_in")
    _count()
    return out


def groupnorm_stats(x: Tensor, groups: int = 32) -> Tensor:
    _need_cuda(x)
    B, H, W, C_ = x.shape
    sums = torch.empty(B, groups, 2, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().tvae_groupnorm_stats(x.data_ptr(), sums.data_ptr(), B, H * W, C_, groups, _stream()),
               "tvae_groupnorm_stats")
    _count(2)
    return sums


def groupnorm_silu(x: Tensor, gamma: Tensor, beta: Tensor, groups: int = 32, eps: float = 1e-5, silu: bool = True,
                   sums: Optional[Tensor] = None) -> Tensor:
    """silu(GroupNorm(x)) on NHWC bf16."""
    _need_cuda(x, gamma, beta)
    assert x.dtype == BF16 and x.is_contiguous()
    B, H, W, C_ = x.shape
    if sums is None:
        sums = groupnorm_stats(x, groups)
    y = torch.empty_like(x)
    g, b = gamma.float().contiguous(), beta.float().contiguous()
    _lib.check(_lib.load().tvae_groupnorm_apply(x.data_ptr(), sums.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(),
                                                B, H * W, C_, groups, eps, 1 if silu else 0, _stream()),
               "tvae_groupnorm_apply")
    _count()
    return y


def row_stats(x: Tensor, w1: Optional[Tensor] = None, mode: Optional[int] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """mode 0 (w1 None): rstd of RMSNorm.  mode 1: (1/(sigma*rms), mu/sigma) of LayerNorm(RMSNorm(x)*w1).
    mode 2: (1/sigma, mu/sigma) of LayerNorm(x*w1) (no RMSNorm in front)."""
    _need_cuda(x, w1)
    assert x.dtype == BF16 and x.is_contiguous()
    C_ = x.shape[-1]
    M = x.numel() // C_
    if mode is None:
        mode = 0 if w1 is None else 1
    a = torch.empty(M, dtype=torch.float32, device=x.device)
    b = torch.empty(M, dtype=torch.float32, device=x.device) if mode != 0 else None
    w1f = None if w1 is None else w1.float().contiguous()
    _lib.check(_lib.load().tvae_row_stats(x.data_ptr(), _ptr(w1f), a.data_ptr(), _ptr(b), M, C_, mode,
                                          _stream()), "tvae_row_stats")
    _count()
    return a, b


def nchw_to_nhwc(x: Tensor, cpad: int) -> Tensor:
    _need_cuda(x)
    x = x.float().contiguous()
    B, C_, H, W = x.shape
    out = torch.empty(B, H, W, cpad, dtype=BF16, device=x.device)
    _lib.check(_lib.load().tvae_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), B, C_, H, W, cpad, _stream()),
               "tvae_nchw_to_nhwc")
    _count()
    return out


def nhwc_to_nchw(x: Tensor, c: Optional[int] = None) -> Tensor:
    _need_cuda(x)
    assert x.dtype == BF16 and x.is_contiguous()
    B, H, W, Cs = x.shape
    c = Cs if c is None else c
    out = torch.empty(B, c, H, W, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().tvae_nhwc_to_nchw(x.data_ptr(), out.data_ptr(), B, c, H, W, Cs, _stream()),
               "tvae_nhwc_to_nchw")
    _count()
    return out


def reparam(mu: Tensor, logvar: Tensor, eps: Tensor, patched: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """Returns (z, mu', logvar') where the primed tensors are clamped when ``patched``."""
    _need_cuda(mu, logvar, eps)
    mu, logvar, eps = mu.float().contiguous(), logvar.float().contiguous(), eps.float().contiguous()
    z = torch.empty_like(mu)
    mu_o = torch.empty_like(mu) if patched else None
    lv_o = torch.empty_like(mu) if patched else None
    _lib.check(_lib.load().tvae_reparam(mu.data_ptr(), logvar.data_ptr(), eps.data_ptr(), z.data_ptr(), _ptr(mu_o),
                                        _ptr(lv_o), mu.numel(), 1 if patched else 0, _stream()), "tvae_reparam")
    _count()
    return z, (mu_o if patched else mu), (lv_o if patched else logvar)


def loss_sums(recon: Tensor, target: Tensor, mu: Tensor, logvar: Tensor, patched: bool,
              clip: Tuple[float, float] = (-30.0, 20.0)) -> Tensor:
    """fp32[4] = (sum |f(recon) - target|, sum KL terms, #non-finite, 0)."""
    _need_cuda(recon, target, mu, logvar)
    recon, target = recon.float().contiguous(), target.float().contiguous()
    mu, logvar = mu.float().contiguous(), logvar.float().contiguous()
    acc = torch.empty(4, dtype=torch.float32, device=recon.device)
    _lib.check(_lib.load().tvae_loss_l1_kl(recon.data_ptr(), target.data_ptr(), mu.data_ptr(), logvar.data_ptr(),
                                           acc.data_ptr(), recon.numel(), mu.numel(), 1 if patched else 0,
                                           float(clip[0]), float(clip[1]), _stream()), "tvae_loss_l1_kl")
    _count(2)
    return acc
