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
# same for the HBM-bound kernels: (name, algorithmic_bytes, start_event, end_event); None = off
PROFILE_HBM = None
# bumped by trainer.FusedAdamW whenever the (flat) parameters change behind autograd's back: invalidates the per-step
# cache of packed weight operands (_autograd._w_pack)
WEIGHT_EPOCH = 0


# SMs left to NCCL while gradient buckets are all-reduced under the backward pass (TVAE_COMM_RESERVED_SMS, default 0:
# full-width grids; see trainer._set_comm_active)
import os as _os
COMM_RESERVED_SMS = int(_os.environ.get("TVAE_COMM_RESERVED_SMS", "0"))
_comm_active = False


def set_comm_active(on: bool) -> None:
    global _comm_active
    if COMM_RESERVED_SMS > 0 and on != _comm_active:
        _lib.load().tvae_set_reserved_sms(COMM_RESERVED_SMS if on else 0)
    _comm_active = on


class _hbm:
    """with _hbm(name, algorithmic_bytes): launch(...) -- CUDA-event timing of one bandwidth-bound launch when
    ``PROFILE_HBM`` is a list (bench.py's HBM roofline table); free otherwise."""
    __slots__ = ("name", "nbytes", "e0")

    def __init__(self, name: str, nbytes: float):
        self.name, self.nbytes, self.e0 = name, nbytes, None

    def __enter__(self):
        if PROFILE_HBM is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE_HBM.append((self.name, float(self.nbytes), self.e0, e1))
        return False


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


def _fill_taps(d, plan: Plan) -> None:
    d.num_phases = plan.num_phases
    for ph, taps in enumerate(plan.phases):
        d.ntaps[ph] = len(taps)
        d.out_p[ph] = plan.out_p[ph]
        d.out_c_off[ph] = plan.out_c_off[ph]
        for i, t in enumerate(taps):
            dt = d.taps[ph][i]
            dt.map, dt.c_off, dt.dw, dt.p, dt.dh, dt.kblocks, dt.wk_off = t.map, t.c_off, t.dw, t.p, t.dh, t.kblocks, t.wk_off


def _exec_k(plan: Plan) -> int:
    """K per output element the kernel really executes (<= the reference's ``algo_k``: nearest-2x + 3x3 runs as four
    phase-specific 2x2 convolutions; zero padding of K is not counted as work)."""
    k = sum(t.kblocks * 64 for t in plan.phases[0])
    return min(k, plan.algo_k) if plan.algo_k else k


def mtgemm(plan: Plan, a0: Tensor, w: Tensor, *, a1: Optional[Tensor] = None, out: Optional[Tensor] = None,
           out_shape: Optional[Sequence[int]] = None, bias: Optional[Tensor] = None, act: int = ACT_NONE,
           residual: Optional[Tensor] = None, row_scale: Optional[Tensor] = None, row_shift: Optional[Tensor] = None,
           col_sum: Optional[Tensor] = None, rope: Optional[Tuple[Tensor, int, int, int, float]] = None,
           out_f32: Optional[Tensor] = None, out_f32_shape: Optional[Sequence[int]] = None, out_n: int = 0,
           act_grad_z: Optional[Tensor] = None, gn_groups: int = 0, dual: bool = False,
           gn_bwd: Optional[Tuple[Tensor, Tensor, Tensor, Tensor, int, float, bool]] = None):
    """Launch ``tvae_mtgemm``.  a0 / a1 / out / residual are NHWC bf16 4-D tensors (flat matrices as
    [1, 1, M, K]); ``w`` is the packed bf16 [N, K_total] weight; ``bias`` fp32 [phases, N] (or [N]).

    ``gn_groups`` > 0: the launch also produces the GroupNorm statistics of its output (per image and group: sum and
    sum of squares, fp64 [B, gn_groups, 2]) -- from the GEMM epilogue when the launch qualifies, see the header -- and
    attaches them to the returned tensor as ``out._gn_sums`` for ``groupnorm_silu(..., sums=...)``.

    Backward fusion: with ``act_grad_z`` (the saved pre-activation, same shape as the output) and ``act`` set, the launch
    returns ``acc * act'(z)`` -- or ``(acc + residual) * act'(z)`` when ``residual`` is given too (GELU, plain views).

    ``dual=True`` (training forward, ``act`` set, plain bias epilogue): returns ``(z, act(z))`` -- the pre-activation the
    backward pass needs and the activation the next layer reads, stored by one launch.

    ``gn_bwd=(x, sums, gamma, beta, groups, eps, silu)`` (input-gradient GEMM whose output dh feeds the backward of
    ``act(GroupNorm(x))``; plain epilogue): the launch also leaves the reduce pass of that backward -- per (image, channel)
    (sum dy, sum dy * xhat), fp32 [B, C, 2] -- as ``out._gnb_part`` for ``groupnorm_bwd(..., part=...)``."""
    _need_cuda(a0, w, a1, out, bias, residual, row_scale, row_shift, col_sum, out_f32)
    assert w.dtype == BF16 and w.is_contiguous() and w.shape[1] == plan.k_total, (w.shape, plan.k_total)
    d = MtGemmDesc()
    _set_view(d.a0, a0, plan.a0_split)
    _set_view(d.a1, a1, plan.a1_split)
    _fill_taps(d, plan)
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
    d.act_grad, d.z = 0, None
    if act_grad_z is not None:
        _need_cuda(act_grad_z)
        assert act_grad_z.dtype == BF16 and act_grad_z.is_contiguous() and act_grad_z.shape == out.shape and act != ACT_NONE
        if residual is None:
            d.act_grad, residual = 1, act_grad_z             # z rides in the staged ("residual") tile
        else:
            d.act_grad, d.z = 2, act_grad_z.data_ptr()
    _set_view(d.res, residual, plan.out_split)
    if residual is not None:
        assert residual.shape == out.shape
    d.w = w.data_ptr()
    d.n_total, d.k_total = n_total, plan.k_total
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
    gn_sums = None
    if gn_groups:
        assert out_f32 is None and out.shape[-1] == n_total
        gn_sums = torch.empty(out.shape[0], gn_groups, 2, dtype=torch.float64, device=a0.device)
    d.gn_sums, d.gn_groups = _ptr(gn_sums), gn_groups
    out_act = None
    if dual:
        assert out_f32 is None and act != ACT_NONE and residual is None and act_grad_z is None and not gn_groups
        out_act = torch.empty_like(out)
    d.out_act = _ptr(out_act)
    gnb_part = gnb_keep = None
    if gn_bwd is not None:
        gx, gs, gg, gb, ggroups, geps, gsilu = gn_bwd
        _need_cuda(gx, gs, gg, gb)
        assert out_f32 is None and gx.dtype == BF16 and gx.is_contiguous() and gx.shape == out.shape and out.shape[-1] == n_total
        assert gs.dtype == torch.float64 and gs.is_contiguous() and tuple(gs.shape) == (out.shape[0], ggroups, 2)
        gnb_keep = (gg.float().contiguous(), gb.float().contiguous())
        gnb_part = torch.empty(out.shape[0], n_total, 2, dtype=torch.float32, device=a0.device)
        d.gnb_x, d.gnb_sums, d.gnb_gamma, d.gnb_beta = gx.data_ptr(), gs.data_ptr(), gnb_keep[0].data_ptr(), gnb_keep[1].data_ptr()
        d.gnb_part, d.gnb_groups, d.gnb_eps, d.gnb_silu = gnb_part.data_ptr(), ggroups, geps, 1 if gsilu else 0
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(_lib.load().tvae_mtgemm(C.byref(d), _stream()), f"tvae_mtgemm[{plan.name}]")
    if PROFILE is not None:
        e1.record()
        o = out if out_f32 is None else out_f32
        n_real = out_n if out_f32 is not None else n_total
        m_out = o.numel() // (o.shape[1] if out_f32 is not None else o.shape[-1])
        tag = f"{plan.name} M={m_out} N={n_total} K={plan.k_total} act={act} res={int(residual is not None)} rs={int(row_scale is not None)} rope={int(rope is not None)}"
        if gn_bwd is not None:
            tag += " +gn_bwd_reduce"
        PROFILE.append((tag, 2.0 * m_out * n_real * plan.algo_k, e0, e1, 2.0 * m_out * n_real * _exec_k(plan)))
    _count()
    if gn_sums is not None:
        out._gn_sums = gn_sums
    if gnb_part is not None:
        out._gnb_part = gnb_part
    if dual:
        return out, out_act
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
        PROFILE.append(("attn_fwd", 4.0 * B * S * S * C_, e0, e1, 4.0 * B * S * S * C_))
    _count()
    return out, lse


def im2col_in(x: Tensor) -> Tensor:
    """NCHW fp32 image [B, 3, H, W] -> bf16 [B*H*W, 64] im2col rows (hi | lo | 1 | 0) for the first convolution."""
    _need_cuda(x)
    x = x.float().contiguous()
    B, Cin, H, W = x.shape
    if Cin != 3:
        raise ValueError(f"conv_in: 3 input channels expected, got {Cin}")
    cols = torch.empty(B * H * W, 64, dtype=BF16, device=x.device)
    with _hbm("im2col_in", x.numel() * 4 + cols.numel() * 2):
        _lib.check(_lib.load().tvae_im2col_in(x.data_ptr(), cols.data_ptr(), B, H, W, _stream()), "tvae_im2col_in")
    _count()
    return cols


def conv_in(x: Tensor, w: Tensor, b: Optional[Tensor], gn_groups: int = 0) -> Tensor:
    """encoder.conv_in: NCHW fp32 image -> NHWC bf16 features (im2col + K = 64 tensor-core GEMM, bias in column 54).
    ``gn_groups``: also take the GroupNorm statistics of the output (``out._gn_sums``, see ``mtgemm``)."""
    _need_cuda(x, w, b)
    from . import _taps as T
    B, _, H, W = x.shape
    co = w.shape[0]
    cols = im2col_in(x)
    w27 = w.detach().float().reshape(co, 27)
    bias = torch.zeros(co, 1, dtype=torch.float32, device=w.device) if b is None else b.detach().float().reshape(co, 1)
    wp = torch.cat([w27, w27, bias, torch.zeros(co, 9, dtype=torch.float32, device=w.device)], dim=1).to(BF16).contiguous()
    # one image per row of the tile grid ([B, 1, H*W, .] views), so that a 128-pixel tile never straddles two images
    out = mtgemm(T.plan_linear(64), cols.view(B, 1, H * W, 64), wp, out_shape=(B, 1, H * W, co), gn_groups=gn_groups)
    o4 = out.view(B, H, W, co)
    if gn_groups:
        o4._gn_sums = out._gn_sums
    return o4


def groupnorm_stats(x: Tensor, groups: int = 32) -> Tensor:
    _need_cuda(x)
    B, H, W, C_ = x.shape
    sums = torch.empty(B, groups, 2, dtype=torch.float64, device=x.device)
    with _hbm("gn_stats", x.numel() * 2):
        _lib.check(_lib.load().tvae_groupnorm_stats(x.data_ptr(), sums.data_ptr(), B, H * W, C_, groups, _stream()),
                   "tvae_groupnorm_stats")
    _count(2)
    return sums


def groupnorm_silu(x: Tensor, gamma: Tensor, beta: Tensor, groups: int = 32, eps: float = 1e-5, silu: bool = True,
                   sums: Optional[Tensor] = None, return_sums: bool = False):
    """silu(GroupNorm(x)) on NHWC bf16.  Without precomputed ``sums`` both passes run in one call
    (``tvae_groupnorm_silu``); ``return_sums`` also hands back the statistics for the backward pass."""
    _need_cuda(x, gamma, beta)
    assert x.dtype == BF16 and x.is_contiguous()
    B, H, W, C_ = x.shape
    y = torch.empty_like(x)
    g, b = gamma.float().contiguous(), beta.float().contiguous()
    if sums is None:
        sums = torch.empty(B, groups, 2, dtype=torch.float64, device=x.device)
        with _hbm("gn_stats + gn_apply_silu", x.numel() * 6):
            _lib.check(_lib.load().tvae_groupnorm_silu(x.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(), sums.data_ptr(),
                                                       B, H * W, C_, groups, eps, 1 if silu else 0, _stream()),
                       "tvae_groupnorm_silu")
        _count(3)
    else:
        assert sums.dtype == torch.float64 and sums.is_contiguous() and tuple(sums.shape) == (B, groups, 2), (sums.dtype, sums.shape)
        with _hbm("gn_apply_silu", x.numel() * 4):
            _lib.check(_lib.load().tvae_groupnorm_apply(x.data_ptr(), sums.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(),
                                                        B, H * W, C_, groups, eps, 1 if silu else 0, _stream()),
                       "tvae_groupnorm_apply")
        _count()
    return (y, sums) if return_sums else y


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
    with _hbm("row_stats", x.numel() * 2):
        _lib.check(_lib.load().tvae_row_stats(x.data_ptr(), _ptr(w1f), a.data_ptr(), _ptr(b), M, C_, mode,
                                              _stream()), "tvae_row_stats")
    _count()
    return a, b


def nchw_to_nhwc(x: Tensor, cpad: int) -> Tensor:
    _need_cuda(x)
    x = x.float().contiguous()
    B, C_, H, W = x.shape
    out = torch.empty(B, H, W, cpad, dtype=BF16, device=x.device)
    with _hbm("nchw_to_nhwc", x.numel() * 4 + out.numel() * 2):
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
    with _hbm("nhwc_to_nchw", x.numel() * 2 + out.numel() * 4):
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
    with _hbm("reparam", mu.numel() * (24 if patched else 16)):
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
    with _hbm("loss_l1_kl", (recon.numel() * 2 + mu.numel() * 2) * 4):
        _lib.check(_lib.load().tvae_loss_l1_kl(recon.data_ptr(), target.data_ptr(), mu.data_ptr(), logvar.data_ptr(),
                                               acc.data_ptr(), recon.numel(), mu.numel(), 1 if patched else 0,
                                               float(clip[0]), float(clip[1]), _stream()), "tvae_loss_l1_kl")
    _count(2)
    return acc


# ------------------------------------------------------------------------------------------------
# backward pass
# ------------------------------------------------------------------------------------------------
def mtgemm_wgrad(plan: Plan, a0: Tensor, dz: Tensor, n_total: int, a1: Optional[Tensor] = None, bias: bool = False,
                 dw_out: Optional[Tensor] = None, db_out: Optional[Tensor] = None):
    """dW [n_total, k_total] fp32 of the forward ``mtgemm(plan, a0, w, a1=a1)`` given dZ (the gradient w.r.t. its
    pre-activation output, same NHWC bf16 layout / view as the forward output).  ``bias=True`` also returns the bias
    gradient fp32 [num_phases, n_total] (per-phase column sums of dZ), produced by the same launch.
    ``dw_out`` / ``db_out``: accumulate into these fp32 buffers (the kernel adds with ``red.global.add``) instead of
    zeroed new ones."""
    _need_cuda(a0, dz, a1)
    d = MtGemmDesc()
    _set_view(d.a0, a0, plan.a0_split)
    _set_view(d.a1, a1, plan.a1_split)
    _set_view(d.out, dz, plan.out_split)
    _fill_taps(d, plan)
    d.n_total, d.k_total = n_total, plan.k_total
    if dw_out is None:
        dw = torch.zeros(n_total, plan.k_total, dtype=torch.float32, device=a0.device)
    else:
        assert dw_out.dtype == torch.float32 and dw_out.is_contiguous() and dw_out.numel() == n_total * plan.k_total \
            and dw_out.device == a0.device
        dw = dw_out
    if bias and db_out is not None:
        assert db_out.dtype == torch.float32 and db_out.is_contiguous() and db_out.numel() == plan.num_phases * n_total \
            and db_out.device == a0.device
        db = db_out.view(plan.num_phases, n_total)
    else:
        db = torch.zeros(plan.num_phases, n_total, dtype=torch.float32, device=a0.device) if bias else None
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if bias:
        _lib.check(_lib.load().tvae_mtgemm_wgrad_bias(C.byref(d), dw.data_ptr(), db.data_ptr(), _stream()),
                   f"tvae_mtgemm_wgrad_bias[{plan.name}]")
    else:
        _lib.check(_lib.load().tvae_mtgemm_wgrad(C.byref(d), dw.data_ptr(), _stream()), f"tvae_mtgemm_wgrad[{plan.name}]")
    if PROFILE is not None:
        e1.record()
        m_out = dz.numel() // dz.shape[-1]
        PROFILE.append((f"wgrad {plan.name} M={m_out} N={n_total} K={plan.k_total}", 2.0 * m_out * n_total * plan.algo_k, e0, e1,
                        2.0 * m_out * n_total * _exec_k(plan)))
    _count(2)
    return (dw, db) if bias else dw


def bias_act_bwd(dy: Tensor, z: Optional[Tensor], act: int, phase_view: bool = False) -> Tuple[Tensor, Tensor]:
    """(dZ, column sums of dZ).  dy / z: bf16 with channels last.  ``phase_view``: dy is [B, 2H, 2W, C] and the sums
    are returned per output phase as fp32 [2, 2, C] (p, q, c)."""
    _need_cuda(dy, z)
    assert dy.dtype == BF16 and dy.is_contiguous()
    dz = torch.empty_like(dy) if act != ACT_NONE else dy
    zp = _ptr(z) if act != ACT_NONE else None
    dzp = dz.data_ptr() if act != ACT_NONE else None
    if phase_view:
        B, H2, W2, Cc = dy.shape
        cs = torch.empty(2, 2 * Cc, dtype=torch.float32, device=dy.device)
        with _hbm("bias_act_bwd" if act != ACT_NONE else "bias_grad (column sums)", dy.numel() * (6 if act != ACT_NONE else 2)):
            _lib.check(_lib.load().tvae_bias_act_bwd_4d(dy.data_ptr(), zp, dzp, cs.data_ptr(), B * (H2 // 2), 2, W2 // 2,
                                                        2 * Cc, act, _stream()), "tvae_bias_act_bwd_4d")
        _count(2)
        return dz, cs.view(2, 2, Cc)
    N = dy.shape[-1]
    M = dy.numel() // N
    cs = torch.empty(N, dtype=torch.float32, device=dy.device)
    with _hbm("bias_act_bwd" if act != ACT_NONE else "bias_grad (column sums)", dy.numel() * (6 if act != ACT_NONE else 2)):
        _lib.check(_lib.load().tvae_bias_act_bwd(dy.data_ptr(), zp, dzp, cs.data_ptr(), M, N, act, _stream()),
                   "tvae_bias_act_bwd")
    _count(2)
    return dz, cs


def act_fwd(z: Tensor, act: int) -> Tensor:
    _need_cuda(z)
    assert z.dtype == BF16 and z.is_contiguous()
    y = torch.empty_like(z)
    with _hbm("act_fwd", z.numel() * 4):
        _lib.check(_lib.load().tvae_act_fwd(z.data_ptr(), y.data_ptr(), z.numel(), act, _stream()), "tvae_act_fwd")
    _count()
    return y


def groupnorm_bwd(x: Tensor, dh: Tensor, sums: Tensor, gamma: Tensor, beta: Tensor, add: Optional[Tensor] = None,
                  groups: int = 32, eps: float = 1e-5, silu: bool = True,
                  part: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """Backward of h = act(GroupNorm(x)): returns (dx [+ add], dgamma, dbeta).  ``part``: the reduce pass's result when the
    GEMM that produced ``dh`` already left it (``mtgemm(..., gn_bwd=...)``) -- only the apply pass runs then."""
    _need_cuda(x, dh, sums, gamma, beta, add, part)
    B, H, W, Cc = x.shape
    assert sums.dtype == torch.float64 and sums.is_contiguous() and tuple(sums.shape) == (B, groups, 2), (sums.dtype, sums.shape)
    g, b = gamma.float().contiguous(), beta.float().contiguous()
    dx = torch.empty_like(x)
    if part is not None:
        assert part.dtype == torch.float32 and part.is_contiguous() and tuple(part.shape) == (B, Cc, 2)
        with _hbm("gn_bwd (apply)", x.numel() * (8 if add is not None else 6)):
            _lib.check(_lib.load().tvae_groupnorm_bwd_apply(x.data_ptr(), dh.data_ptr(), _ptr(add), sums.data_ptr(),
                                                            g.data_ptr(), b.data_ptr(), part.data_ptr(), dx.data_ptr(), B,
                                                            H * W, Cc, groups, eps, 1 if silu else 0, _stream()),
                       "tvae_groupnorm_bwd_apply")
        _count(1)
    else:
        part = torch.empty(B, Cc, 2, dtype=torch.float32, device=x.device)
        with _hbm("gn_bwd (reduce + apply)", x.numel() * (12 if add is not None else 10)):
            _lib.check(_lib.load().tvae_groupnorm_bwd(x.data_ptr(), dh.data_ptr(), _ptr(add), sums.data_ptr(), g.data_ptr(),
                                                      b.data_ptr(), part.data_ptr(), dx.data_ptr(), B, H * W, Cc, groups, eps,
                                                      1 if silu else 0, _stream()), "tvae_groupnorm_bwd")
        _count(3)
    red = part.sum(0)          # [C, 2]: tiny (B x C) reduction of the per-image partials
    return dx, red[:, 1], red[:, 0]          # strided views (consumers: GradSink / autograd accept them)


def token_norm_fwd(x: Tensor, w: Tensor, mode: int) -> Tensor:
    _need_cuda(x, w)
    assert x.dtype == BF16 and x.is_contiguous()
    Cc = x.shape[-1]
    y = torch.empty_like(x)
    wf = w.float().contiguous()
    with _hbm("token_norm_fwd", x.numel() * 4):
        _lib.check(_lib.load().tvae_token_norm_fwd(x.data_ptr(), wf.data_ptr(), y.data_ptr(), x.numel() // Cc, Cc, mode,
                                                   _stream()), "tvae_token_norm_fwd")
    _count()
    return y


def token_norm_bwd(x: Tensor, w: Tensor, dy: Tensor, add: Optional[Tensor], mode: int) -> Tuple[Tensor, Tensor]:
    _need_cuda(x, w, dy, add)
    Cc = x.shape[-1]
    dx = torch.empty_like(x)
    dw = torch.empty(Cc, dtype=torch.float32, device=x.device)
    wf = w.float().contiguous()
    with _hbm("token_norm_bwd", x.numel() * (8 if add is not None else 6)):
        _lib.check(_lib.load().tvae_token_norm_bwd(x.data_ptr(), wf.data_ptr(), dy.data_ptr(), _ptr(add), dx.data_ptr(),
                                                   dw.data_ptr(), x.numel() // Cc, Cc, mode, _stream()), "tvae_token_norm_bwd")
    _count(2)
    return dx, dw


def attn_bwd(qkv: Tensor, out: Tensor, dout: Tensor, lse: Tensor, rope_tab: Tensor, B: int, S: int, C_: int, H: int,
             W: int, q_scale: float) -> Tensor:
    """Gradient w.r.t. the PRE-RoPE, pre-scale q | k | v projection output, bf16 [B, S, 3C]."""
    _need_cuda(qkv, out, dout, lse, rope_tab)
    lib = _lib.load()
    dev = qkv.device
    delta = torch.empty(B, C_ // 64, S, dtype=torch.float32, device=dev)
    slices = lib.tvae_attn_bwd_dq_slices(S)      # 2: ordered (bit-reproducible) dQ accumulation, even / odd steps
    dq_acc = torch.empty(slices, B, S, C_, dtype=torch.float32, device=dev)
    dqkv = torch.empty(B, S, 3 * C_, dtype=BF16, device=dev)
    dout = dout.contiguous()
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(lib.tvae_attn_delta(out.data_ptr(), dout.data_ptr(), delta.data_ptr(), B, S, C_, _stream()), "tvae_attn_delta")
    _lib.check(lib.tvae_attn_bwd(qkv.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq_acc.data_ptr(),
                                 dqkv.data_ptr(), B, S, C_, slices, _stream()), "tvae_attn_bwd")
    _lib.check(lib.tvae_rope_bwd(dq_acc.data_ptr(), dqkv.data_ptr(), rope_tab.data_ptr(), B * S, C_, H, W, q_scale,
                                 slices, _stream()), "tvae_rope_bwd")
    if PROFILE is not None:
        e1.record()
        PROFILE.append(("attn_bwd", 10.0 * B * S * S * C_, e0, e1, 10.0 * B * S * S * C_))
    _count(4)
    return dqkv


def conv_in_wgrad(x: Tensor, dy: Tensor) -> Tuple[Tensor, Tensor]:
    """(dW [Co, 3, 3, 3], dbias [Co]) of encoder.conv_in from dY (NHWC bf16): one tcgen05 wgrad launch on the im2col rows."""
    _need_cuda(x, dy)
    from . import _taps as T
    B, _, H, W = x.shape
    co = dy.shape[-1]
    cols = im2col_in(x)
    dwp = mtgemm_wgrad(T.plan_linear(64), cols.view(1, 1, B * H * W, 64), dy.reshape(1, 1, B * H * W, co), co)
    dw = (dwp[:, :27] + dwp[:, 27:54]).reshape(co, 3, 3, 3).contiguous()
    return dw, dwp[:, 54].contiguous()


def loss_bwd(recon: Tensor, target: Tensor, mu: Tensor, logvar: Tensor, scal: Tensor, patched: bool,
             clip: Tuple[float, float]) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda(recon, target, mu, logvar, scal)
    drecon, dmu, dlv = torch.empty_like(recon), torch.empty_like(mu), torch.empty_like(logvar)
    with _hbm("loss_bwd", (recon.numel() * 3 + mu.numel() * 4) * 4):
        _lib.check(_lib.load().tvae_loss_bwd(recon.data_ptr(), target.data_ptr(), mu.data_ptr(), logvar.data_ptr(),
                                             scal.data_ptr(), drecon.data_ptr(), dmu.data_ptr(), dlv.data_ptr(), recon.numel(),
                                             mu.numel(), 1 if patched else 0, float(clip[0]), float(clip[1]), _stream()),
                   "tvae_loss_bwd")
    _count(2)
    return drecon, dmu, dlv


def latent_bwd(mu: Tensor, logvar: Tensor, eps: Tensor, dz: Optional[Tensor], dmu_ret: Optional[Tensor],
               dlv_ret: Optional[Tensor], patched: bool) -> Tuple[Tensor, Tensor]:
    _need_cuda(mu, logvar, eps, dz, dmu_ret, dlv_ret)
    dmu, dlv = torch.empty_like(mu), torch.empty_like(mu)
    c = lambda t: None if t is None else t.float().contiguous()
    dz, dmu_ret, dlv_ret = c(dz), c(dmu_ret), c(dlv_ret)
    _lib.check(_lib.load().tvae_latent_bwd(mu.data_ptr(), logvar.data_ptr(), eps.data_ptr(), _ptr(dz), _ptr(dmu_ret),
                                           _ptr(dlv_ret), dmu.data_ptr(), dlv.data_ptr(), mu.numel(), 1 if patched else 0,
                                           _stream()), "tvae_latent_bwd")
    _count()
    return dmu, dlv


def dwconv3x3(u: Tensor, w9c: Tensor, bias: Optional[Tensor], flip: bool = False, add_input: bool = True) -> Tensor:
    """y = [u +] depthwise3x3(u; w9c) [+ bias] on NHWC bf16; w9c fp32 [9, C] (tap-major).  ``flip``: the spatially
    flipped taps (input gradient of the same layer)."""
    _need_cuda(u, w9c, bias)
    assert u.dtype == BF16 and u.is_contiguous() and u.dim() == 4
    B, H, W, C_ = u.shape
    assert w9c.dtype == torch.float32 and w9c.is_contiguous() and tuple(w9c.shape) == (9, C_)
    y = torch.empty_like(u)
    with _hbm("dwconv3x3", u.numel() * 4):
        _lib.check(_lib.load().tvae_dwconv3x3(u.data_ptr(), w9c.data_ptr(), _ptr(bias), y.data_ptr(), B, H, W, C_,
                                              1 if flip else 0, 1 if add_input else 0, _stream()), "tvae_dwconv3x3")
    _count()
    return y


def dwconv3x3_wgrad(u: Tensor, dy: Tensor, bias: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
    """(dw fp32 [9, C], db fp32 [C]) of y = u + depthwise3x3(u) + b given dy (both NHWC bf16)."""
    _need_cuda(u, dy)
    assert u.dtype == BF16 and dy.dtype == BF16 and u.is_contiguous() and dy.is_contiguous() and u.shape == dy.shape
    B, H, W, C_ = u.shape
    dw = torch.empty(9, C_, dtype=torch.float32, device=u.device)
    db = torch.empty(C_, dtype=torch.float32, device=u.device) if bias else None
    with _hbm("dwconv3x3_wgrad", u.numel() * 4):
        _lib.check(_lib.load().tvae_dwconv3x3_wgrad(u.data_ptr(), dy.data_ptr(), dw.data_ptr(), _ptr(db), B, H, W, C_,
                                                    _stream()), "tvae_dwconv3x3_wgrad")
    _count(3)
    return dw, db


def weight_pack(w: Tensor, fwd: bool = True, dgrad: bool = False) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """fp32 parameter in the reference layout -- nn.Linear [A, B] or nn.Conv2d 3x3 [A, B, 3, 3] -- to the bf16 operands of
    the tensor-core kernels: forward [A, T*B] (== ``_taps.pack_conv3x3`` + cast) and / or input-gradient [B, T*A]
    (== ``_taps.pack_conv3x3_dgrad`` + cast), one coalesced pass each (``tvae_weight_pack``)."""
    _need_cuda(w)
    assert w.dtype == torch.float32 and w.is_contiguous() and w.dim() in (2, 4), (w.dtype, w.shape)
    A, B_ = w.shape[0], w.shape[1]
    T_ = 1 if w.dim() == 2 else w.shape[2] * w.shape[3]
    wf = torch.empty(A, T_ * B_, dtype=BF16, device=w.device) if fwd else None
    wd = torch.empty(B_, T_ * A, dtype=BF16, device=w.device) if dgrad else None
    with _hbm("weight_pack", w.numel() * (4 + 2 * (int(fwd) + int(dgrad)))):
        _lib.check(_lib.load().tvae_weight_pack(w.data_ptr(), _ptr(wf), _ptr(wd), A, B_, T_, _stream()), "tvae_weight_pack")
    _count()
    return wf, wd


def wgrad_unpack(g: Tensor, shape: Sequence[int], accumulate_into: Optional[Tensor] = None) -> Tensor:
    """Packed weight gradient fp32 [A, 9*B] (``mtgemm_wgrad`` of a 3x3 convolution) -> parameter layout [A, B, 3, 3];
    ``accumulate_into``: fp32 contiguous tensor of that layout (a ``.grad`` slot) the gradient is ADDED to."""
    _need_cuda(g, accumulate_into)
    A, B_ = int(shape[0]), int(shape[1])
    assert g.dtype == torch.float32 and g.is_contiguous() and tuple(g.shape) == (A, 9 * B_) and tuple(shape[2:]) == (3, 3)
    if accumulate_into is not None:
        out = accumulate_into
        assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == g.numel()
    else:
        out = torch.empty(A, B_, 3, 3, dtype=torch.float32, device=g.device)
    with _hbm("wgrad_unpack", g.numel() * (12 if accumulate_into is not None else 8)):
        _lib.check(_lib.load().tvae_wgrad_unpack(g.data_ptr(), out.data_ptr(), A, B_, 9, int(accumulate_into is not None),
                                                 _stream()), "tvae_wgrad_unpack")
    _count()
    return out


def _ptr3(ts: Sequence[Tensor]):
    import ctypes
    return (ctypes.c_void_p * 3)(*[t.data_ptr() for t in ts])


def fold_qkv(w3: Sequence[Tensor], g3: Sequence[Tensor], b3: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
    """([Wq g_q; Wk g_k; Wv g_v] fp32 [3C, C], [Wq b_q; Wk b_k; Wv b_v] fp32 [3C]) in one launch (``tvae_fold_qkv``; the
    torch expression is ``_taps.fold_qkv_affine``)."""
    C_ = w3[0].shape[0]
    for t in (*w3, *g3, *b3):
        _need_cuda(t)
        assert t.dtype == torch.float32 and t.is_contiguous(), (t.dtype, t.shape)
    wg = torch.empty(3 * C_, C_, dtype=torch.float32, device=w3[0].device)
    bg = torch.empty(3 * C_, dtype=torch.float32, device=w3[0].device)
    _lib.check(_lib.load().tvae_fold_qkv(_ptr3(w3), _ptr3(g3), _ptr3(b3), wg.data_ptr(), bg.data_ptr(), C_, _stream()),
               "tvae_fold_qkv")
    _count()
    return wg, bg


def fold_qkv_bwd(w3, g3, b3, dwg: Tensor, dbg: Tensor, accumulate_into=None):
    """Gradients of the nine inputs of ``fold_qkv`` (three lists: dw [C, C], dg [C], db [C]).  ``accumulate_into`` =
    (dw3, dg3, db3): fp32 contiguous ``.grad`` slots the gradients are ADDED to (returned as they are)."""
    C_ = w3[0].shape[0]
    dwg, dbg = dwg.float().contiguous(), dbg.float().contiguous()
    dev = dwg.device
    if accumulate_into is not None:
        dw, dg, db = (list(t) for t in accumulate_into)
        for t in (*dw, *dg, *db):
            _need_cuda(t)
            assert t.dtype == torch.float32 and t.is_contiguous()
        assert all(t.numel() == C_ * C_ for t in dw) and all(t.numel() == C_ for t in (*dg, *db))
    else:
        dw = [torch.empty(C_, C_, dtype=torch.float32, device=dev) for _ in range(3)]
        dg = [torch.empty(C_, dtype=torch.float32, device=dev) for _ in range(3)]
        db = [torch.empty(C_, dtype=torch.float32, device=dev) for _ in range(3)]
    _lib.check(_lib.load().tvae_fold_qkv_bwd(_ptr3(w3), _ptr3(g3), _ptr3(b3), dwg.data_ptr(), dbg.data_ptr(), _ptr3(dw),
                                             _ptr3(dg), _ptr3(db), C_, int(accumulate_into is not None), _stream()),
               "tvae_fold_qkv_bwd")
    _count()
    return dw, dg, db


def upconv1_pack(t: Tensor, backward: bool = False) -> Tensor:
    """Upsample conv1 weight [O, I, 3, 3] -> phase-packed [O, 16*I] (``_taps.pack_upsample_conv1``), or -- backward -- the
    packed gradient [O, 16*I] -> [O, I, 3, 3]."""
    _need_cuda(t)
    t = t.float().contiguous()
    if backward:
        O, I = t.shape[0], t.shape[1] // 16
        out = torch.empty(O, I, 3, 3, dtype=torch.float32, device=t.device)
    else:
        O, I = t.shape[0], t.shape[1]
        out = torch.empty(O, 16 * I, dtype=torch.float32, device=t.device)
    _lib.check(_lib.load().tvae_upconv1_pack(t.data_ptr(), out.data_ptr(), O, I, int(backward), _stream()), "tvae_upconv1_pack")
    _count()
    return out


SUMSQ_BLOCKS = 1024          # TVAE_SUMSQ_BLOCKS


def grad_sumsq(g: Tensor, partials: Tensor) -> None:
    """partials fp64 [SUMSQ_BLOCKS] = fixed-order per-block sums of g^2 (flat fp32 or bf16 buffer, length % 8 == 0)."""
    _need_cuda(g, partials)
    assert partials.dtype == torch.float64 and partials.numel() == SUMSQ_BLOCKS and g.dtype in (torch.float32, BF16)
    with _hbm("grad_sumsq (grad norm)", g.numel() * g.element_size()):
        _lib.check(_lib.load().tvae_grad_sumsq(g.data_ptr(), int(g.dtype == BF16), g.numel(), partials.data_ptr(), _stream()),
                   "tvae_grad_sumsq")
    _count()


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, partials: Tensor, state: Tensor, lr_base: float,
               warmup_steps: int = 0, betas=(0.9, 0.95), eps: float = 1e-8, weight_decay: float = 0.0,
               max_norm: float = 1.0, grad_scale: float = 1.0) -> None:
    """Fused clip + AdamW + warm-up + non-finite skip over flat buffers; every decision is taken on the device from
    ``partials`` (``grad_sumsq``) and ``state`` (fp32 [8], see include/transvae_sm100.h)."""
    _need_cuda(p, g, m, v, partials, state)
    assert state.dtype == torch.float32 and state.numel() == 8 and g.numel() == p.numel()
    with _hbm("adamw", p.numel() * (24 + g.element_size())):
        _lib.check(_lib.load().tvae_adamw_step(p.data_ptr(), g.data_ptr(), int(g.dtype == BF16), m.data_ptr(), v.data_ptr(),
                                               p.numel(), partials.data_ptr(), state.data_ptr(), float(lr_base),
                                               int(warmup_steps), float(betas[0]), float(betas[1]), float(eps),
                                               float(weight_decay), float(max_norm or 0.0), float(grad_scale), _stream()),
                   "tvae_adamw_step")
    _count(2)


def cast_f32_bf16(src: Tensor, dst: Tensor) -> None:
    _need_cuda(src, dst)
    assert src.dtype == torch.float32 and dst.dtype == BF16 and src.numel() == dst.numel()
    with _hbm("grad cast fp32->bf16", src.numel() * 6):
        _lib.check(_lib.load().tvae_cast_f32_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "tvae_cast_f32_bf16")
    _count()


def multi_tensor_add(dsts: Sequence[Tensor], srcs: Sequence[Tensor]) -> None:
    """dst[i] += src[i] for many small fp32 tensors in one launch per 96 tensors (``tvae_multi_tensor_add``).  A source
    may be a 1-D strided view (e.g. one column of a [C, 2] matrix)."""
    n = len(dsts)
    if n == 0:
        return
    _need_cuda(*dsts, *srcs)
    keep, strides = [], []
    for d, s in zip(dsts, srcs):
        assert d.dtype == torch.float32 and d.is_contiguous() and d.numel() == s.numel(), (d.shape, s.shape, d.dtype)
        if s.dtype == torch.float32 and s.dim() == 1 and s.stride(0) >= 1:
            strides.append(s.stride(0))
        else:
            if s.dtype != torch.float32 or not s.is_contiguous():
                s = s.float().contiguous()
            strides.append(1)
        keep.append(s)
    PtrArr, IntArr = C.c_void_p * n, C.c_int32 * n
    da, sa = PtrArr(*[d.data_ptr() for d in dsts]), PtrArr(*[s.data_ptr() for s in keep])
    na, st = IntArr(*[d.numel() for d in dsts]), IntArr(*strides)
    with _hbm("multi_tensor_add", sum(d.numel() for d in dsts) * 12):
        _lib.check(_lib.load().tvae_multi_tensor_add(da, sa, na, st, n, _stream()), "tvae_multi_tensor_add")
    _count((n + 95) // 96)


METRIC_MODES = {None: 0, "none": 0, "identity": 0, "clamp": 1, "sigmoid": 2}


def metrics_sums(recon: Tensor, target: Tensor, mode: str = "clamp") -> Tensor:
    """fp32 [B, 4] = per image (sum squared error, sum |error|, sum of the 11x11 SSIM map, 0) of f(recon) vs target."""
    _need_cuda(recon, target)
    recon, target = recon.float().contiguous(), target.float().contiguous()
    assert recon.shape == target.shape and recon.dim() == 4
    B, C_, H, W = recon.shape
    acc = torch.empty(B, 4, dtype=torch.float32, device=recon.device)
    with _hbm("metrics (psnr/ssim)", recon.numel() * 8):
        _lib.check(_lib.load().tvae_metrics(recon.data_ptr(), target.data_ptr(), acc.data_ptr(), B, C_, H, W,
                                            METRIC_MODES[mode], _stream()), "tvae_metrics")
    _count()
    return acc
