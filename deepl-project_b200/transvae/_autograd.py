"""Training path: block-level ``torch.autograd.Function``s whose forward AND backward are hand-written kernels.

Granularity is one Function per reference module (ResBlock, Downsample, Upsample, the attention half and the FFN half
of a TransVAEBlock, the first / last convolutions, the latent heads, reparameterisation, loss) so that residual adds,
bias gradients and activation derivatives stay fused inside the kernels and autograd only chains block to block.

Weights enter as fp32 *packed* tensors produced by differentiable torch ops on the reference-layout parameters
(``_taps.pack_*`` -- weight-side plumbing); the Functions return fp32 gradients for them (tcgen05 wgrad kernel) and
autograd maps those back to the parameters.  Activations and their gradients are NHWC bf16.

Unlike the inference path (which folds RMSNorm / LayerNorm into GEMM epilogues), the training path materialises the
normalised token tensors with ``tvae_token_norm_fwd`` so that the projections stay plain GEMMs in the backward pass.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from . import _taps as T
from . import ops
from ._lib import ACT_GELU, ACT_NONE, ACT_SILU

Tensor = torch.Tensor
BF16 = torch.bfloat16
Fn = torch.autograd.Function


def _bf(w: Tensor) -> Tensor:
    return w.detach().to(BF16).contiguous()


def _tr(wp: Tensor, taps: int) -> Tensor:
    """Packed forward weight [N, taps*C] -> dgrad weight [C, taps*N] (bf16)."""
    n = wp.shape[0]
    c = wp.shape[1] // taps
    return wp.detach().view(n, taps, c).permute(2, 1, 0).reshape(c, taps * n).to(BF16).contiguous()


def _f32(t: Tensor) -> Tensor:
    return t.detach().float().contiguous()


# Reference-layout weights ([out, in] linears, [out, in, 3, 3] convolutions) -> bf16 kernel operands and back.  fp32
# contiguous weights take one coalesced ``tvae_weight_pack`` / ``tvae_wgrad_unpack`` pass; anything else the torch route
# (TVAE_TORCH_PACK=1 forces it: A/B switch for the bench).
_TORCH_PACK = os.environ.get("TVAE_TORCH_PACK", "0") == "1"
# TVAE_DIRECT_GRAD=0: no per-step operand cache and no wgrad accumulation straight into the trainer's flat gradient slots
_DIRECT = os.environ.get("TVAE_DIRECT_GRAD", "1") != "0"


class FoldQkvFn(torch.autograd.Function):
    """``_taps.fold_qkv_affine`` as one kernel each way (``ops.fold_qkv`` / ``ops.fold_qkv_bwd``)."""

    @staticmethod
    def forward(ctx, wq, wk, wv, gq, bq, gk, bk, gv, bv):
        ts = [t.detach().float().contiguous() for t in (wq, wk, wv, gq, bq, gk, bk, gv, bv)]
        ctx.save_for_backward(*ts)
        ctx.params = (wq, wk, wv, gq, bq, gk, bk, gv, bv)
        return ops.fold_qkv(ts[0:3], ts[3::2], ts[4::2])

    @staticmethod
    def backward(ctx, dwg, dbg):
        ts = ctx.saved_tensors
        # all nine inputs are trainer-owned leaves: the kernel adds straight into their flat-buffer gradient slots (autograd's
        # AccumulateGrad was nine tiny add_ launches per attention block, 234 per micro-step of the large model)
        ps = ctx.params
        slots = [_direct_slot(p, p.numel(), False) for p in ps]
        if all(s is not None for s in slots):
            ops.fold_qkv_bwd(ts[0:3], ts[3::2], ts[4::2], dwg, dbg, accumulate_into=(slots[0:3], slots[3::2], slots[4::2]))
            for p in ps:
                _fire_hooks(p)
            return (None,) * 9
        dw, dg, db = ops.fold_qkv_bwd(ts[0:3], ts[3::2], ts[4::2], dwg, dbg)
        return dw[0], dw[1], dw[2], dg[0], db[0], dg[1], db[1], dg[2], db[2]


class UpConv1PackFn(torch.autograd.Function):
    """``_taps.pack_upsample_conv1`` as one kernel each way (``ops.upconv1_pack``)."""

    @staticmethod
    def forward(ctx, w):
        return ops.upconv1_pack(w.detach())

    @staticmethod
    def backward(ctx, dpacked):
        return ops.upconv1_pack(dpacked, backward=True)


def fold_qkv_affine(wq, wk, wv, gq, bq, gk, bk, gv, bv):
    """Device tensors: the kernels; anything else (CPU unit tests of the host logic): the torch expression."""
    if wq.is_cuda:
        return FoldQkvFn.apply(wq, wk, wv, gq, bq, gk, bk, gv, bv)
    from . import _taps
    return _taps.fold_qkv_affine(wq, wk, wv, gq, bq, gk, bk, gv, bv)


def pack_upsample_conv1(w):
    if w.is_cuda:
        return UpConv1PackFn.apply(w)
    from . import _taps
    return _taps.pack_upsample_conv1(w)


_PACKS: dict = {}          # id(parameter) -> (signature, (forward operand, input-gradient operand))


def clear_pack_cache() -> None:
    _PACKS.clear()


def _packable(w: Tensor) -> bool:
    if _TORCH_PACK or w.dtype != torch.float32 or not w.is_contiguous():
        return False
    if w.dim() == 2:
        return w.shape[0] % 4 == 0 and w.shape[1] % 4 == 0
    return w.dim() == 4 and tuple(w.shape[2:]) == (3, 3) and w.shape[0] % 2 == 0 and w.shape[1] % 2 == 0


def _w_pack(w: Tensor):
    """(forward operand [out, taps*in], input-gradient operand [in, taps*out]), both bf16, from ONE read of ``w``: the
    second is kept for the backward pass."""
    # parameters owned by trainer.GradBuckets only change in the optimizer step (which bumps ops.WEIGHT_EPOCH): with
    # gradient accumulation their operands are packed once per step, not once per micro-step
    cached = _DIRECT and getattr(w, "_tvae_direct_grad", False) and w.is_leaf
    if cached:
        sig = (ops.WEIGHT_EPOCH, w._version, w.data_ptr(), tuple(w.shape))
        hit = _PACKS.get(id(w))
        if hit is not None and hit[0] == sig:
            return hit[1]
    w0, w = w, w.detach()
    if w.dim() == 4 and w.shape[2] == 1 and w.shape[3] == 1:        # 1x1 convolution = matrix
        w = w.reshape(w.shape[0], w.shape[1])
    if _packable(w):
        out = ops.weight_pack(w, fwd=True, dgrad=True)
        if cached:
            _PACKS[id(w0)] = (sig, out)
        return out
    if w.dim() == 4:
        return _bf(T.pack_conv3x3(w)), _bf(T.pack_conv3x3_dgrad(w))
    return _bf(w), _bf(w.t())


def _w_ungrad(gp: Tensor, w: Tensor) -> Optional[Tensor]:
    """packed fp32 gradient [out, taps*in] -> the layout of ``w`` (None when it went straight into ``w``'s gradient slot)."""
    if w.dim() == 2:
        return gp
    if gp.dtype == torch.float32 and gp.is_contiguous():
        slot = _direct_slot(w, w.numel(), False)
        if slot is not None:            # re-layout and accumulation into the flat gradient slot in one pass
            ops.wgrad_unpack(gp, w.shape, accumulate_into=slot)
            _fire_hooks(w)
            return None
        return ops.wgrad_unpack(gp, w.shape)
    return gp.view(w.shape[0], 3, 3, w.shape[1]).permute(0, 3, 1, 2)


def _fire_hooks(p: Tensor) -> None:
    """Run the post-accumulate-grad hooks of ``p`` by hand (the bucket all-reduce trigger of trainer.GradBuckets): its
    gradient was accumulated into the flat-buffer slot by a kernel, autograd received ``None`` for it."""
    for hook in list((getattr(p, "_post_accumulate_grad_hooks", None) or {}).values()):
        hook(p)


def _wgrad_b(plan, a0: Tensor, dz: Tensor, n_total: int, a1: Optional[Tensor] = None, w: Optional[Tensor] = None,
             b: Optional[Tensor] = None):
    """(dW, dbias) of a single-phase plan in one launch: the bias gradient (column sums of dZ) rides on the wgrad
    kernel's tensor-core pass instead of a separate sweep over dZ.

    ``w`` / ``b``: the weight / bias the gradients belong to.  When one is a leaf whose ``.grad`` is a slot of the
    trainer's flat gradient buffer (``trainer.GradBuckets`` marks those parameters; weights must be matrix-shaped), the
    kernel accumulates straight into that slot -- no zero-filled temporary, no ``grad += dW`` pass -- that gradient is
    returned as None and the parameter's post-accumulate hooks (the bucket all-reduce trigger) are run here, since
    autograd has nothing left to accumulate."""
    wslot = _direct_slot(w, n_total * plan.k_total, True) if w is not None else None
    bslot = _direct_slot(b, n_total, False) if b is not None and plan.num_phases == 1 else None
    dw, db = ops.mtgemm_wgrad(plan, a0, dz, n_total, a1=a1, bias=True, dw_out=wslot, db_out=bslot)
    for p, slot in ((w, wslot), (b, bslot)):
        if slot is not None:
            _fire_hooks(p)
    return (None if wslot is not None else dw), (None if bslot is not None else db[0])


def _direct_slot(p: Tensor, numel: int, matrix: bool) -> Optional[Tensor]:
    ok = _DIRECT and getattr(p, "_tvae_direct_grad", False) and p.is_leaf and p.requires_grad
    g = p.grad if ok else None
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.numel() != numel or p.numel() != numel:
        return None
    if matrix and not (p.dim() == 2 or (p.dim() == 4 and p.shape[2] == 1 and p.shape[3] == 1)):
        return None
    return g


class GradSink:
    """Gradient accumulation of the SMALL parameters (norm weights / biases, convolution biases of multi-phase plans,
    3x3 weights after ``tvae_wgrad_unpack``) without autograd's one-tiny-kernel-per-parameter AccumulateGrad: a backward
    Function hands (parameter, gradient) pairs to ``add`` -- which answers None, so autograd has nothing to accumulate --
    and ``flush`` adds all of them into their flat-buffer slots with ONE ``tvae_multi_tensor_add`` launch per 96 tensors
    and then runs the parameters' post-accumulate hooks (bucket all-reduce triggers).  Parameters that are not owned by
    ``trainer.GradBuckets`` pass through untouched."""
    BIG = 1 << 18          # larger tensors take their own (bandwidth-efficient) add

    def __init__(self):
        self._dst, self._src, self._params = [], [], []

    def add(self, p: Optional[Tensor], g: Optional[Tensor]) -> Optional[Tensor]:
        if p is None or g is None:
            return g
        slot = _direct_slot(p, p.numel(), False)
        if slot is None or g.numel() != p.numel() or not g.is_cuda:
            return g
        if g.numel() > self.BIG:
            slot.view(-1).add_(g.reshape(-1))
            _fire_hooks(p)
            return None
        self._dst.append(slot.view(-1))
        self._src.append(g if g.dim() == 1 else g.reshape(-1))
        self._params.append(p)
        return None

    def flush(self) -> None:
        if not self._dst:
            return
        dst, src, params = self._dst, self._src, self._params
        self._dst, self._src, self._params = [], [], []
        ops.multi_tensor_add(dst, src)
        for p in params:
            _fire_hooks(p)


GRAD_SINK = GradSink()


def _flat(x: Tensor) -> Tensor:
    return x.reshape(1, 1, -1, x.shape[-1])


# ------------------------------------------------------------------------------------------------
class ResBlockFn(Fn):
    """x + conv2(silu(GN2(conv1(silu(GN1(x))))))  (blocks.py:58-68)."""

    @staticmethod
    def forward(ctx, x, x_sums, g1, b1, w1, c1b, g2, b2, w2, c2b):
        # w1 / w2: conv weights in the reference layout [C, C, 3, 3]; x_sums: GroupNorm statistics of x when the producer
        # left them (the previous ResBlock's conv2 epilogue), else None.  Returns (out, statistics of out).
        B, H, W, C = x.shape
        plan = T.plan_conv3x3(C)
        h0, s1 = ops.groupnorm_silu(x, g1, b1, sums=x_sums, return_sums=True)
        (w1f, w1d), (w2f, w2d) = _w_pack(w1), _w_pack(w2)
        h1 = ops.mtgemm(plan, h0, w1f, out_shape=(B, H, W, C), bias=_f32(c1b), gn_groups=32)   # + GN2's statistics
        h2, s2 = ops.groupnorm_silu(h1, g2, b2, sums=h1._gn_sums, return_sums=True)
        out = ops.mtgemm(plan, h2, w2f, out_shape=(B, H, W, C), bias=_f32(c2b), residual=x, gn_groups=32)
        out_sums = out._gn_sums
        ctx.save_for_backward(x, s1, h0, h1, s2, h2, g1, b1, w1, g2, b2, w2, w1d, w2d, c1b, c2b)
        ctx.mark_non_differentiable(out_sums)
        return out, out_sums

    @staticmethod
    def backward(ctx, dout, _dsums):
        x, s1, h0, h1, s2, h2, g1, b1, w1, g2, b2, w2, w1d, w2d, c1b, c2b = ctx.saved_tensors
        dout = dout.contiguous()
        B, H, W, C = x.shape
        dplan = T.plan_conv3x3_dgrad(C)
        dw2, dc2b = _wgrad_b(T.plan_conv3x3(C), h2, dout, C, b=c2b)
        # the input-gradient GEMMs leave the reduce pass of the GroupNorm backward that consumes their output
        dh2 = ops.mtgemm(dplan, dout, w2d, out_shape=(B, H, W, C), gn_bwd=(h1, s2, g2, b2, 32, 1e-5, True))
        dh1, dg2, db2 = ops.groupnorm_bwd(h1, dh2, s2, g2, b2, part=dh2._gnb_part)
        dw1, dc1b = _wgrad_b(T.plan_conv3x3(C), h0, dh1, C, b=c1b)
        dh0 = ops.mtgemm(dplan, dh1, w1d, out_shape=(B, H, W, C), gn_bwd=(x, s1, g1, b1, 32, 1e-5, True))
        dx, dg1, db1 = ops.groupnorm_bwd(x, dh0, s1, g1, b1, add=dout, part=dh0._gnb_part)
        S = GRAD_SINK
        out = (dx, None, S.add(g1, dg1), S.add(b1, db1), S.add(w1, _w_ungrad(dw1, w1)), S.add(c1b, dc1b), S.add(g2, dg2),
               S.add(b2, db2), S.add(w2, _w_ungrad(dw2, w2)), S.add(c2b, dc2b))
        S.flush()
        return out


class ResBlockScFn(Fn):
    """conv2(silu(GN2(conv1(silu(GN1(x)))))) + shortcut(x) with a convolutional shortcut (blocks.py:40-46, :68; in != out
    channels).  ``w2s`` = packed [Cout, 9*Cout + k*k*Cin] (conv2 | shortcut), ``b2s`` = conv2.bias + shortcut.bias: conv2
    and the shortcut share one accumulator (``_taps.plan_resblock_conv2``)."""

    @staticmethod
    def forward(ctx, x, x_sums, g1, b1, w1, c1b, g2, b2, w2s, b2s, k):
        B, H, W, Cin = x.shape
        Cout = w1.shape[0]
        h0, s1 = ops.groupnorm_silu(x, g1, b1, sums=x_sums, return_sums=True)
        w1f, w1d = _w_pack(w1)
        h1 = ops.mtgemm(T.plan_conv3x3(Cin), h0, w1f, out_shape=(B, H, W, Cout), bias=_f32(c1b), gn_groups=32)
        h2, s2 = ops.groupnorm_silu(h1, g2, b2, sums=h1._gn_sums, return_sums=True)
        out = ops.mtgemm(T.plan_resblock_conv2(Cout, Cin, k), h2, _bf(w2s), a1=x, out_shape=(B, H, W, Cout), bias=_f32(b2s),
                         gn_groups=32)
        out_sums = out._gn_sums
        ctx.save_for_backward(x, s1, h0, h1, s2, h2, g1, b1, w1, g2, b2, w2s, w1d, c1b)
        ctx.k = k
        ctx.mark_non_differentiable(out_sums)
        return out, out_sums

    @staticmethod
    def backward(ctx, dout, _dsums):
        x, s1, h0, h1, s2, h2, g1, b1, w1, g2, b2, w2s, w1d, c1b = ctx.saved_tensors
        dout = dout.contiguous()
        k = ctx.k
        B, H, W, Cin = x.shape
        Cout = w1.shape[0]
        dw2s, db2s = _wgrad_b(T.plan_resblock_conv2(Cout, Cin, k), h2, dout, Cout, a1=x)
        dh2 = ops.mtgemm(T.plan_conv3x3_dgrad(Cout), dout, _tr(w2s[:, :9 * Cout], 9), out_shape=(B, H, W, Cout),
                         gn_bwd=(h1, s2, g2, b2, 32, 1e-5, True))
        dx_sc = ops.mtgemm(T.plan_conv_kxk_dgrad(Cout, k), dout, _tr(w2s[:, 9 * Cout:], k * k), out_shape=(B, H, W, Cin))
        dh1, dg2, db2 = ops.groupnorm_bwd(h1, dh2, s2, g2, b2, part=dh2._gnb_part)
        dw1, dc1b = _wgrad_b(T.plan_conv3x3(Cin), h0, dh1, Cout, b=c1b)
        dh0 = ops.mtgemm(T.plan_conv3x3_dgrad(Cout), dh1, w1d, out_shape=(B, H, W, Cin),
                         gn_bwd=(x, s1, g1, b1, 32, 1e-5, True))
        dx, dg1, db1 = ops.groupnorm_bwd(x, dh0, s1, g1, b1, add=dx_sc, part=dh0._gnb_part)
        S = GRAD_SINK
        out = (dx, None, S.add(g1, dg1), S.add(b1, db1), S.add(w1, _w_ungrad(dw1, w1)), S.add(c1b, dc1b), S.add(g2, dg2),
               S.add(b2, db2), dw2s, db2s, None)
        S.flush()
        return out


class FfnDwFn(Fn):
    """x + proj_out(u + dwconv3x3(u)), u = gelu(proj_in(RMSNorm(x; w2)))  (conv.py:42-50, 79-105: conv_type='depthwise').
    ``wdw``: the depthwise weight as fp32 [9, hid] (tap-major)."""

    @staticmethod
    def forward(ctx, x, w2n, win, bin_, wdw, bdw, wout, bout):
        B, H, W, C = x.shape
        M = B * H * W
        hid = win.shape[0]
        (win_f, win_d), (wout_f, wout_d) = _w_pack(win), _w_pack(wout)
        xn = ops.token_norm_fwd(x, w2n, 0)
        z_in, u = ops.mtgemm(T.plan_linear(C), _flat(xn), win_f, out_shape=(1, 1, M, hid), bias=_f32(bin_), act=ACT_GELU,
                             dual=True)
        wdw_c = _f32(wdw)
        u2 = ops.dwconv3x3(u.view(B, H, W, hid), wdw_c, _f32(bdw), flip=False, add_input=True)
        out = ops.mtgemm(T.plan_linear(hid), _flat(u2), wout_f, out_shape=(1, 1, M, C), bias=_f32(bout), residual=_flat(x))
        ctx.save_for_backward(x, w2n, xn, z_in, u, u2, wdw_c, win_d, wout_d, win, wout, bin_, bout)
        return out.view(B, H, W, C)

    @staticmethod
    def backward(ctx, dout):
        x, w2n, xn, z_in, u, u2, wdw_c, win_d, wout_d, win, wout, bin_, bout = ctx.saved_tensors
        dout = dout.contiguous()
        B, H, W, C = x.shape
        M = B * H * W
        hid = win.shape[0]
        df = _flat(dout)
        dwout, dbout = _wgrad_b(T.plan_linear(hid), _flat(u2), df, C, w=wout, b=bout)
        du2 = ops.mtgemm(T.plan_linear(C), df, wout_d, out_shape=(1, 1, M, hid)).view(B, H, W, hid)
        dwdw, dbdw = ops.dwconv3x3_wgrad(u.view(B, H, W, hid), du2)
        du = ops.dwconv3x3(du2, wdw_c, None, flip=True, add_input=True)            # du2 + dwconv^T(du2)
        dzin, _ = ops.bias_act_bwd(du.view(M, hid), z_in.view(M, hid), ACT_GELU)
        dzin = dzin.view(1, 1, M, hid)
        dwin, dbin = _wgrad_b(T.plan_linear(C), _flat(xn), dzin, hid, w=win, b=bin_)
        dxn = ops.mtgemm(T.plan_linear(hid), dzin, win_d, out_shape=(1, 1, M, C))
        dx, dw2n = ops.token_norm_bwd(x, w2n, dxn.view(B, H, W, C), dout, 0)
        dw2n = GRAD_SINK.add(w2n, dw2n)
        GRAD_SINK.flush()
        return dx, dw2n, dwin, dbin, dwdw, dbdw, dwout, dbout


class DownsampleFn(Fn):
    """conv3x3 s2 (silu(conv3x3(x))) + conv1x1(pixel_unshuffle(x))  (upsample.py:55-66)."""

    @staticmethod
    def forward(ctx, x, w0p, b0, wdp, bd):
        B, H, W, C = x.shape
        N = wdp.shape[0]
        dc = wdp.shape[1] == 13 * C            # [N, 9C] without the DC path (use_dc_path=False)
        z0, y = ops.mtgemm(T.plan_conv3x3(C), x, _bf(w0p), out_shape=(B, H, W, C), bias=_f32(b0), act=ACT_SILU, dual=True)
        out = ops.mtgemm(T.plan_downsample(C, with_dc=dc), y, _bf(wdp), a1=x if dc else None,
                         out_shape=(B, H // 2, W // 2, N), bias=_f32(bd))
        ctx.save_for_backward(x, z0, y, w0p, wdp)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, z0, y, w0p, wdp = ctx.saved_tensors
        dout = dout.contiguous()
        B, H, W, C = x.shape
        N = wdp.shape[0]
        dc = wdp.shape[1] == 13 * C
        dwd, dbd = _wgrad_b(T.plan_downsample(C, with_dc=dc), y, dout, N, a1=x if dc else None)
        # dZ0 = dY * silu'(Z0) leaves the input-gradient GEMM directly (act_grad epilogue); its bias gradient rides on wgrad
        dz0 = ops.mtgemm(T.plan_downsample_dgrad_main(C, N), dout, _tr(wdp[:, :9 * C], 9), out_shape=(B, H, W, C),
                         act=ACT_SILU, act_grad_z=z0)
        dx_dc = ops.mtgemm(T.plan_downsample_dgrad_dc(C, N), dout, _tr(wdp[:, 9 * C:], 4), out_shape=(B, H, W, C)) if dc else None
        dw0, db0 = _wgrad_b(T.plan_conv3x3(C), x, dz0, C)
        dx = ops.mtgemm(T.plan_conv3x3_dgrad(C), dz0, _tr(w0p, 9), out_shape=(B, H, W, C), residual=dx_dc)
        return dx, dw0, db0, dwd, dbd


class UpsampleFn(Fn):
    """conv3x3(silu(conv3x3(nearest2x(x)))) + pixel_shuffle(conv1x1(x))  (upsample.py:116-126)."""

    @staticmethod
    def forward(ctx, x, w1p, b1, w2p, b2p):
        B, H, W, Ci = x.shape
        Co = w1p.shape[0]
        b1e = _f32(b1).unsqueeze(0).expand(4, -1).contiguous()
        dc = w2p.shape[1] == 9 * Co + 4 * Ci   # [Co, 9 Co] without the DC path (use_dc_path=False)
        z1, y = ops.mtgemm(T.plan_upsample_conv1(Ci, Co), x, _bf(w1p), out_shape=(B, 2 * H, 2 * W, Co), bias=b1e,
                           act=ACT_SILU, dual=True)
        out = ops.mtgemm(T.plan_upsample_conv2(Co, Ci, with_dc=dc), y, _bf(w2p), a1=x if dc else None,
                         out_shape=(B, 2 * H, 2 * W, Co), bias=_f32(b2p))
        ctx.save_for_backward(x, z1, y, w1p, w2p)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, z1, y, w1p, w2p = ctx.saved_tensors
        dout = dout.contiguous()
        B, H, W, Ci = x.shape
        Co = w1p.shape[0]
        dc = w2p.shape[1] == 9 * Co + 4 * Ci
        dw2, db2p = ops.mtgemm_wgrad(T.plan_upsample_conv2(Co, Ci, with_dc=dc), y, dout, Co, a1=x if dc else None,
                                     bias=True)                                                # db per output phase [4, Co]
        dz1 = ops.mtgemm(T.plan_conv3x3_dgrad(Co), dout, _tr(w2p[:, :9 * Co], 9), out_shape=(B, 2 * H, 2 * W, Co),
                         act=ACT_SILU, act_grad_z=z1)
        dx_dc = ops.mtgemm(T.plan_upsample_dc_dgrad(Co), dout, _tr(w2p[:, 9 * Co:], 4), out_shape=(B, H, W, Ci)) if dc else None
        dw1, db1p = ops.mtgemm_wgrad(T.plan_upsample_conv1(Ci, Co), x, dz1, Co, bias=True)
        db1 = db1p.sum(0)                      # the forward bias is shared by the four output phases
        dx = ops.mtgemm(T.plan_upsample_conv1_dgrad(Ci, Co), dz1, _tr(w1p, 16), out_shape=(B, H, W, Ci), residual=dx_dc)
        return dx, dw1, db1, dw2, db2p


class AttnFn(Fn):
    """x + proj(SDPA(rope(q), rope(k), v)), q|k|v = [Wq g_q; Wk g_k; Wv g_v] LNhat(RMSNorm(x; w1)) + [Wq b_q; ...]
    (blocks.py:146, attention.py:65-104).  ``wqkv`` / ``bqkv`` already contain the LayerNorm affines."""

    @staticmethod
    def forward(ctx, x, w1, wqkv, bqkv, wproj, bproj, rope_tab, scale):
        B, H, W, C = x.shape
        S = H * W
        xh = ops.token_norm_fwd(x, w1, 1)
        rope = (rope_tab, C, H, W, scale * math.log2(math.e))
        (wqkv_f, wqkv_d), (wproj_f, wproj_d) = _w_pack(wqkv), _w_pack(wproj)
        qkv = ops.mtgemm(T.plan_linear(C), _flat(xh), wqkv_f, out_shape=(1, 1, B * S, 3 * C), bias=_f32(bqkv), rope=rope)
        o, lse = ops.attn_fwd(qkv.view(B, S, 3 * C), B, S, C, need_lse=True)
        out = ops.mtgemm(T.plan_linear(C), _flat(o), wproj_f, out_shape=(1, 1, B * S, C), bias=_f32(bproj),
                         residual=_flat(x))
        ctx.save_for_backward(x, w1, xh, qkv, o, lse, wqkv_d, wproj_d, rope_tab, wproj, bproj)
        ctx.scale = scale
        return out.view(B, H, W, C)

    @staticmethod
    def backward(ctx, dout):
        x, w1, xh, qkv, o, lse, wqkv_d, wproj_d, rope_tab, wproj, bproj = ctx.saved_tensors
        dout = dout.contiguous()
        B, H, W, C = x.shape
        S = H * W
        df = _flat(dout)
        dwp, dbp = _wgrad_b(T.plan_linear(C), _flat(o), df, C, w=wproj, b=bproj)
        do = ops.mtgemm(T.plan_linear(C), df, wproj_d, out_shape=(1, 1, B * S, C))
        dqkv = ops.attn_bwd(qkv.view(B, S, 3 * C), o, do.view(B, S, C), lse, rope_tab, B, S, C, H, W, ctx.scale)
        dq = _flat(dqkv)
        dwq, dbq = _wgrad_b(T.plan_linear(C), _flat(xh), dq, 3 * C)
        dxh = ops.mtgemm(T.plan_linear(3 * C), dq, wqkv_d, out_shape=(1, 1, B * S, C))
        dx, dw1 = ops.token_norm_bwd(x, w1, dxh.view(B, H, W, C), dout, 1)
        dw1 = GRAD_SINK.add(w1, dw1)
        GRAD_SINK.flush()
        return dx, dw1, dwq, dbq, dwp, dbp, None, None


class FfnFn(Fn):
    """x + proj_out(u + conv(u)), u = gelu(proj_in(RMSNorm(x; w2)))  (blocks.py:149, conv.py:79-105)."""

    @staticmethod
    def forward(ctx, x, w2n, win, bin_, wc0, bc0, wc2, bc2, wc4, bc4, wout, bout):
        # win / wc0 / wc4 / wout: [out, in] matrices; wc2: the 3x3 conv weight in the reference layout [mid, mid, 3, 3]
        B, H, W, C = x.shape
        M = B * H * W
        hid, mid = win.shape[0], wc0.shape[0]
        (win_f, win_d), (wc0_f, wc0_d), (wc2_f, wc2_d) = _w_pack(win), _w_pack(wc0), _w_pack(wc2)
        (wc4_f, wc4_d), (wout_f, wout_d) = _w_pack(wc4), _w_pack(wout)
        xn = ops.token_norm_fwd(x, w2n, 0)
        z_in, u = ops.mtgemm(T.plan_linear(C), _flat(xn), win_f, out_shape=(1, 1, M, hid), bias=_f32(bin_), act=ACT_GELU,
                             dual=True)
        z0, t0 = ops.mtgemm(T.plan_linear(hid), u, wc0_f, out_shape=(1, 1, M, mid), bias=_f32(bc0), act=ACT_GELU, dual=True)
        z2, t2 = ops.mtgemm(T.plan_conv3x3(mid), t0.view(B, H, W, mid), wc2_f, out_shape=(B, H, W, mid), bias=_f32(bc2),
                            act=ACT_GELU, dual=True)
        u2 = ops.mtgemm(T.plan_linear(mid), _flat(t2), wc4_f, out_shape=(1, 1, M, hid), bias=_f32(bc4), residual=u)
        out = ops.mtgemm(T.plan_linear(hid), u2, wout_f, out_shape=(1, 1, M, C), bias=_f32(bout), residual=_flat(x))
        ctx.save_for_backward(x, w2n, xn, z_in, u, z0, t0, z2, t2, u2, wc2, win_d, wc0_d, wc2_d, wc4_d, wout_d, win, wc0, wc4, wout,
                              bin_, bc0, bc2, bc4, bout)
        ctx.dims = (hid, mid)
        return out.view(B, H, W, C)

    @staticmethod
    def backward(ctx, dout):
        x, w2n, xn, z_in, u, z0, t0, z2, t2, u2, wc2, win_d, wc0_d, wc2_d, wc4_d, wout_d, win, wc0, wc4, wout, bin_, bc0, bc2, bc4, bout = ctx.saved_tensors
        dout = dout.contiguous()
        B, H, W, C = x.shape
        M = B * H * W
        hid, mid = ctx.dims
        df = _flat(dout)
        dwout, dbout = _wgrad_b(T.plan_linear(hid), u2, df, C, w=wout, b=bout)
        du2 = ops.mtgemm(T.plan_linear(C), df, wout_d, out_shape=(1, 1, M, hid))
        dwc4, dbc4 = _wgrad_b(T.plan_linear(mid), _flat(t2), du2, hid, w=wc4, b=bc4)
        # every dZ = dY * gelu'(Z) of the block is produced by the epilogue of the GEMM that computes dY (act_grad), every
        # bias gradient by the wgrad launch that consumes dZ: no separate pass over the [M, 4C] / [M, C] gradients
        dz2 = ops.mtgemm(T.plan_linear(hid), du2, wc4_d, out_shape=(1, 1, M, mid), act=ACT_GELU,
                         act_grad_z=z2.view(1, 1, M, mid))
        dz2i = dz2.view(B, H, W, mid)
        dwc2, dbc2 = _wgrad_b(T.plan_conv3x3(mid), t0.view(B, H, W, mid), dz2i, mid, b=bc2)
        dz0 = ops.mtgemm(T.plan_conv3x3_dgrad(mid), dz2i, wc2_d, out_shape=(B, H, W, mid), act=ACT_GELU,
                         act_grad_z=z0.view(B, H, W, mid)).view(1, 1, M, mid)
        dwc0, dbc0 = _wgrad_b(T.plan_linear(hid), u, dz0, mid, w=wc0, b=bc0)
        dzin = ops.mtgemm(T.plan_linear(mid), dz0, wc0_d, out_shape=(1, 1, M, hid), residual=du2,
                          act=ACT_GELU, act_grad_z=z_in)
        dwin, dbin = _wgrad_b(T.plan_linear(C), _flat(xn), dzin, hid, w=win, b=bin_)
        dxn = ops.mtgemm(T.plan_linear(hid), dzin, win_d, out_shape=(1, 1, M, C))
        dx, dw2n = ops.token_norm_bwd(x, w2n, dxn.view(B, H, W, C), dout, 0)
        def _as(g, w):                      # 1x1 conv weights arrive as [out, in, 1, 1]
            return g if g is None else g.view_as(w)
        S = GRAD_SINK
        out = (dx, S.add(w2n, dw2n), dwin, dbin, _as(dwc0, wc0), dbc0, S.add(wc2, _w_ungrad(dwc2, wc2)), dbc2, _as(dwc4, wc4),
               dbc4, dwout, dbout)
        S.flush()
        return out


class ConvInFn(Fn):
    """encoder.conv_in (encoder.py:111): NCHW fp32 image -> NHWC bf16 features; no input gradient."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x)
        return ops.conv_in(x, w.detach(), None if b is None else b.detach())

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dw, db = ops.conv_in_wgrad(x, dy.contiguous())
        return None, dw, db


class Conv3x3Fn(Fn):
    """Plain 3x3 conv on an NHWC bf16 input whose channel count is already padded to 64 (decoder.conv_in)."""

    @staticmethod
    def forward(ctx, xn, wp, b):
        B, H, W, C = xn.shape
        out = ops.mtgemm(T.plan_conv3x3(C), xn, _bf(wp), out_shape=(B, H, W, wp.shape[0]), bias=_f32(b))
        ctx.save_for_backward(xn, wp)
        return out

    @staticmethod
    def backward(ctx, dout):
        xn, wp = ctx.saved_tensors
        dout = dout.contiguous()
        B, H, W, C = xn.shape
        N = wp.shape[0]
        dw, db = _wgrad_b(T.plan_conv3x3(C), xn, dout, N)
        dx = ops.mtgemm(T.plan_conv3x3_dgrad(N), dout, _tr(wp, 9), out_shape=(B, H, W, C)) if ctx.needs_input_grad[0] else None
        return dx, dw, db


class HeadFn(Fn):
    """3x3 conv to a few fp32 NCHW channels (conv_mu | conv_logvar fused, transvae.py:182-183; decoder.conv_out,
    decoder.py:130).  ``wp`` / ``b`` are padded to 64 output rows; the first ``n_out`` are real."""

    @staticmethod
    def forward(ctx, h, wp, b, n_out):
        B, H, W, C = h.shape
        out = ops.mtgemm(T.plan_conv3x3(C), h, _bf(wp), bias=_f32(b), out_f32_shape=(B, n_out, H, W), out_n=n_out)
        ctx.save_for_backward(h, wp)
        ctx.n_out = n_out
        return out

    @staticmethod
    def backward(ctx, dout):
        h, wp = ctx.saved_tensors
        B, H, W, C = h.shape
        npad = wp.shape[0]
        dz = ops.nchw_to_nhwc(dout.float().contiguous(), npad)      # zero-padded channels
        dw, db = _wgrad_b(T.plan_conv3x3(C), h, dz, npad)
        dh = ops.mtgemm(T.plan_conv3x3_dgrad(npad), dz, _tr(wp, 9), out_shape=(B, H, W, C))
        return dh, dw, db, None


class GroupNormSilu(Fn):
    """silu(GroupNorm(x)) standalone (decoder.norm_out, decoder.py:128-129)."""

    @staticmethod
    def forward(ctx, x, g, b, silu, x_sums=None):
        y, s = ops.groupnorm_silu(x, g, b, silu=silu, sums=x_sums, return_sums=True)
        ctx.save_for_backward(x, s, g, b)
        ctx.silu = silu
        return y

    @staticmethod
    def backward(ctx, dh):
        x, s, g, b = ctx.saved_tensors
        dx, dg, db = ops.groupnorm_bwd(x, dh.contiguous(), s, g, b, silu=ctx.silu)
        dg, db = GRAD_SINK.add(g, dg), GRAD_SINK.add(b, db)
        GRAD_SINK.flush()
        return dx, dg, db, None, None


class RmsNormFn(Fn):
    """y = x / sqrt(mean(x^2) + eps) * w over the last axis of a bf16 [..., C] tensor (blocks.py:168-204): the stand-alone
    RMSNorm of the reference API; inside the blocks the norm is folded into the consuming GEMM (AttnFn / FfnFn)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return ops.token_norm_fwd(x, w, 0)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dw = ops.token_norm_bwd(x, w, dy.contiguous(), None, 0)
        return dx, dw.to(w.dtype)


class NchwToNhwc(Fn):
    @staticmethod
    def forward(ctx, x, cpad):
        ctx.c = x.shape[1]
        return ops.nchw_to_nhwc(x, cpad)

    @staticmethod
    def backward(ctx, dy):
        return ops.nhwc_to_nchw(dy.contiguous(), ctx.c), None


class NhwcToNchw(Fn):
    @staticmethod
    def forward(ctx, x, c):
        ctx.cs = x.shape[-1]
        return ops.nhwc_to_nchw(x, c)

    @staticmethod
    def backward(ctx, dy):
        return ops.nchw_to_nhwc(dy.contiguous(), ctx.cs), None


class Reparam(Fn):
    """(z, mu', logvar') = reparameterise(mu, logvar; eps)  (transvae.py:186-199, patched :186-196 / :244-245)."""

    @staticmethod
    def forward(ctx, mu, logvar, eps, patched):
        z, mu_c, lv_c = ops.reparam(mu, logvar, eps, patched)
        ctx.save_for_backward(mu, logvar, eps)
        ctx.patched = patched
        if not patched:
            mu_c, lv_c = mu.clone(), logvar.clone()
        return z, mu_c, lv_c

    @staticmethod
    def backward(ctx, dz, dmu_c, dlv_c):
        mu, logvar, eps = ctx.saved_tensors
        dmu, dlv = ops.latent_bwd(mu.float().contiguous(), logvar.float().contiguous(), eps.float().contiguous(), dz, dmu_c,
                                  dlv_c, ctx.patched)
        return dmu, dlv, None, None


class LossSums(Fn):
    """(sum |f(recon) - target|, sum KL terms, #non-finite)  (vae_loss.py:83, :94-95; patched :80-104)."""

    @staticmethod
    def forward(ctx, recon, target, mu, logvar, patched, clip_lo, clip_hi):
        acc = ops.loss_sums(recon, target, mu, logvar, patched, (clip_lo, clip_hi))
        ctx.save_for_backward(recon, target, mu, logvar)
        ctx.cfg = (patched, clip_lo, clip_hi)
        ctx.mark_non_differentiable(acc[2])
        return acc[0], acc[1], acc[2]

    @staticmethod
    def backward(ctx, g_l1, g_kl, _g_bad):
        recon, target, mu, logvar = ctx.saved_tensors
        patched, lo, hi = ctx.cfg
        scal = torch.stack([g_l1.float().reshape(()), g_kl.float().reshape(())]).contiguous()
        dr, dmu, dlv = ops.loss_bwd(recon.float().contiguous(), target.float().contiguous(), mu.float().contiguous(),
                                    logvar.float().contiguous(), scal, patched, (lo, hi))
        return dr, None, dmu, dlv, None, None, None
