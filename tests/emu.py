"""CPU emulator of tvae_mtgemm's addressing (test infrastructure).

Interprets a ``Plan`` exactly like the CUDA kernel does -- 5-D pixel views, per-tap coordinate offsets, zero
fill out of bounds, per-phase output placement -- but in fp32 torch on the CPU, so the host-side tap tables and
weight packing can be checked against F.conv2d without a GPU.
"""
from __future__ import annotations

import torch


def view5(t: torch.Tensor, split: bool) -> torch.Tensor:
    """NHWC [B, H, W, C] -> [B, Hv, P, Wv, Cq]."""
    B, H, W, C = t.shape
    if split:
        return t.reshape(B, H // 2, 2, W // 2, 2 * C)
    return t.reshape(B, H, 1, W, C)


def fetch(v5: torch.Tensor, c_off: int, kc: int, dw: int, p: int, dh: int) -> torch.Tensor:
    B, Hv, _, Wv, _ = v5.shape
    src = v5[:, :, p, :, c_off:c_off + kc]
    out = torch.zeros(B, Hv, Wv, kc, dtype=v5.dtype)
    h0, h1 = max(0, -dh), min(Hv, Hv - dh)
    w0, w1 = max(0, -dw), min(Wv, Wv - dw)
    if h1 > h0 and w1 > w0:
        out[:, h0:h1, w0:w1] = src[:, h0 + dh:h1 + dh, w0 + dw:w1 + dw]
    return out


def emulate(plan, a0, a1, wp, out_shape, bias=None):
    """a0/a1: NHWC fp32; wp: [N, K_total]; out_shape: NHWC shape of the output tensor; bias [phases, N]."""
    v0 = view5(a0, plan.a0_split)
    v1 = view5(a1, plan.a1_split) if a1 is not None else None
    out = torch.zeros(out_shape, dtype=torch.float32)
    o5 = view5(out, plan.out_split)
    N = wp.shape[0]
    for ph, taps in enumerate(plan.phases):
        acc = None
        for t in taps:
            src = v1 if t.map else v0
            kc = t.kblocks * 64
            a = fetch(src, t.c_off, kc, t.dw, t.p, t.dh)
            y = a @ wp[:, t.wk_off:t.wk_off + kc].t()
            acc = y if acc is None else acc + y
        if bias is not None:
            acc = acc + bias[ph]
        o5[:, :, plan.out_p[ph], :, plan.out_c_off[ph]:plan.out_c_off[ph] + N] = acc
    return out


def emulate_wgrad(plan, a0, a1, dz, n_total):
    """dW_packed [n_total, k_total] exactly as tvae_mtgemm_wgrad computes it (fp32)."""
    v0 = view5(a0, plan.a0_split)
    v1 = view5(a1, plan.a1_split) if a1 is not None else None
    z5 = view5(dz, plan.out_split)
    dw = torch.zeros(n_total, plan.k_total)
    for ph, taps in enumerate(plan.phases):
        g = z5[:, :, plan.out_p[ph], :, plan.out_c_off[ph]:plan.out_c_off[ph] + n_total]   # [B, Hv, Wv, N]
        g2 = g.reshape(-1, n_total)
        for t in taps:
            src = v1 if t.map else v0
            kc = t.kblocks * 64
            a = fetch(src, t.c_off, kc, t.dw, t.p, t.dh).reshape(-1, kc)
            dw[:, t.wk_off:t.wk_off + kc] += g2.t() @ a
    return dw
