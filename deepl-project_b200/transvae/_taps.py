"""Host-side planning for the fused multi-tap implicit GEMM (``tvae_mtgemm``).

Every convolution / linear / pixel-(un)shuffle / nearest-upsample of the reference is expressed as a list of
*taps* over "pixel views" of NHWC tensors plus one packed ``[N, K_total]`` weight matrix.  This module builds the
tap tables (``Plan``) and packs reference-layout weights (OIHW / [out, in]) into that matrix.  It is pure
Python / torch and has no device dependency, so the CPU test-suite checks it against ``F.conv2d`` with an
emulator of the kernel's addressing (tests/emu.py).

Phase views (see include/transvae_sm100.h): a tensor [B, H, W, C] viewed as (2C, W/2, 2, H/2, B), i.e. element
(b, 2h+p, 2w+q, c) is channel q*C+c of pixel (b, h, w) in phase p.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch

Tensor = torch.Tensor


@dataclass(frozen=True)
class TapSpec:
    map: int       # 0 / 1: which A view
    c_off: int     # channel offset inside the view (q*C for phase views)
    dw: int        # pixel offsets in view coordinates
    p: int         # phase coordinate (0 for plain views)
    dh: int
    kblocks: int   # number of 64-channel K blocks
    wk_off: int    # column offset in the packed weight matrix


@dataclass
class Plan:
    phases: List[List[TapSpec]]
    out_p: List[int]
    out_c_off: List[int]
    a0_split: bool = False
    a1_split: bool = False
    out_split: bool = False
    k_total: int = 0
    name: str = ""
    algo_k: int = 0   # K per output element at the REFERENCE's cost (2*M*N*algo_k = algorithmic FLOPs)

    @property
    def num_phases(self) -> int:
        return len(self.phases)


def _kb(c: int) -> int:
    if c % 64:
        raise ValueError(f"channel count {c} must be a multiple of 64 for the tensor-core path")
    return c // 64


# ----------------------------------------------------------------------------------------------
# plans
# ----------------------------------------------------------------------------------------------
def plan_linear(k: int) -> Plan:
    """nn.Linear / 1x1 conv on a flat [M, K] token matrix."""
    return Plan([[TapSpec(0, 0, 0, 0, 0, _kb(k), 0)]], [0], [0], k_total=k, name="linear", algo_k=k)


def plan_conv3x3(c: int) -> Plan:
    """3x3, stride 1, pad 1 (blocks.py:34,37; conv.py:57; upsample.py:34,97; heads; conv_out)."""
    taps = [TapSpec(0, 0, dx - 1, 0, dy - 1, _kb(c), (dy * 3 + dx) * c) for dy in range(3) for dx in range(3)]
    return Plan([taps], [0], [0], k_total=9 * c, name="conv3x3", algo_k=9 * c)


def plan_resblock_conv2(cmid: int, cin: int, k: int) -> Plan:
    """conv2 of a ResBlock whose shortcut is a convolution (blocks.py:40-46, :68 ``h + self.shortcut(x)``): the nine taps
    of conv2 over map 0 (the normalised activation, cmid channels) and the k x k taps of the shortcut over map 1 (the
    block input, cin channels) accumulate into ONE accumulator; weights [cout, 9*cmid + k*k*cin]."""
    taps = [TapSpec(0, 0, dx - 1, 0, dy - 1, _kb(cmid), (dy * 3 + dx) * cmid) for dy in range(3) for dx in range(3)]
    r = k // 2
    for dy in range(k):
        for dx in range(k):
            taps.append(TapSpec(1, 0, dx - r, 0, dy - r, _kb(cin), 9 * cmid + (dy * k + dx) * cin))
    return Plan([taps], [0], [0], k_total=9 * cmid + k * k * cin, name=f"resblock_conv2_sc{k}", algo_k=9 * cmid + k * k * cin)


def pack_resblock_conv2(w2: Tensor, ws: Tensor) -> Tensor:
    """[O, 9*Cmid + k*k*Cin]: conv2 taps then the shortcut's (1x1 or 3x3) taps."""
    sc = pack_conv3x3(ws) if ws.shape[-1] == 3 else pack_conv1x1(ws)
    return torch.cat([pack_conv3x3(w2), sc], dim=1)


def plan_conv_kxk_dgrad(n: int, k: int) -> Plan:
    """Input gradient of a k x k (k = 1 or 3), stride-1, same-padded convolution with ``n`` output channels."""
    if k == 3:
        return plan_conv3x3_dgrad(n)
    return Plan([[TapSpec(0, 0, 0, 0, 0, _kb(n), 0)]], [0], [0], k_total=n, name="conv1x1_dgrad", algo_k=n)


def _s2(d: int) -> Tuple[int, int]:
    """Stride-2, pad-1 tap d in {0,1,2} -> (phase, offset) in the 2x phase view: row 2h+d-1."""
    return ((d - 1) % 2, (d - 1) // 2)


def plan_downsample(c: int, with_dc: bool = True) -> Plan:
    """Downsample (upsample.py:55-66): conv3x3 s2 over map 0 (= silu(conv3x3(x))) + 1x1 conv over
    pixel_unshuffle(x) (map 1), both as phase views of [B, H, W, c]; one accumulator, plain output.
    ``with_dc=False``: the main path only (use_dc_path=False, upsample.py:40-42)."""
    taps = []
    for dy in range(3):
        p, dh = _s2(dy)
        for dx in range(3):
            q, dw = _s2(dx)
            taps.append(TapSpec(0, q * c, dw, p, dh, _kb(c), (dy * 3 + dx) * c))
    if not with_dc:
        return Plan([taps], [0], [0], a0_split=True, k_total=9 * c, name="downsample", algo_k=9 * c)
    for i in range(2):
        for j in range(2):
            taps.append(TapSpec(1, j * c, 0, i, 0, _kb(c), (9 + i * 2 + j) * c))
    return Plan([taps], [0], [0], a0_split=True, a1_split=True, k_total=13 * c, name="downsample", algo_k=13 * c)


# nearest-2x followed by 3x3 pad 1: output row 2h+py reads upsampled rows 2h+py+dy-1, i.e. source rows
# h + floor((py+dy-1)/2).  Taps that hit the same source row are summed into one weight.
_UP_ROWS = {0: [(-1, (0,)), (0, (1, 2))], 1: [(0, (0, 1)), (1, (2,))]}


def plan_upsample_conv1(cin: int, cout: int) -> Plan:
    """First conv of Upsample (upsample.py:94-95): nearest2x + conv3x3 == four phase-specific 2x2 convs on the
    low-resolution input (2.25x fewer MACs).  Output is the phase view of [B, 2H, 2W, cout]."""
    phases, out_p, out_c = [], [], []
    wk = 0
    for py in range(2):
        for px in range(2):
            taps = []
            for dh, _ in _UP_ROWS[py]:
                for dw, _ in _UP_ROWS[px]:
                    taps.append(TapSpec(0, 0, dw, 0, dh, _kb(cin), wk))
                    wk += cin
            phases.append(taps)
            out_p.append(py)
            out_c.append(px * cout)
    return Plan(phases, out_p, out_c, out_split=True, k_total=16 * cin, name="upsample_conv1", algo_k=9 * cin)


def plan_upsample_conv2(cmid: int, cin: int, with_dc: bool = True) -> Plan:
    """Second conv of Upsample + DC path (upsample.py:97, 116-126): 3x3 over map 0 (phase view of the
    [B, 2H, 2W, cmid] intermediate) evaluated per output phase, plus pixel_shuffle(conv1x1(x)) which for output
    phase (i, j) is a 1x1 conv over the low-res input x (map 1) with the rows c*4+i*2+j of dc_conv."""
    phases, out_p, out_c = [], [], []
    for py in range(2):
        for px in range(2):
            taps = []
            for dy in range(3):
                sy = py + dy - 1
                for dx in range(3):
                    sx = px + dx - 1
                    taps.append(TapSpec(0, (sx % 2) * cmid, sx // 2, sy % 2, sy // 2, _kb(cmid), (dy * 3 + dx) * cmid))
            if with_dc:
                taps.append(TapSpec(1, 0, 0, 0, 0, _kb(cin), 9 * cmid + (py * 2 + px) * cin))
            phases.append(taps)
            out_p.append(py)
            out_c.append(px * cmid)
    return Plan(phases, out_p, out_c, a0_split=True, out_split=True,
                k_total=9 * cmid + (4 * cin if with_dc else 0), name="upsample_conv2",
                algo_k=9 * cmid + (cin if with_dc else 0))


# ----------------------------------------------------------------------------------------------
# weight packing (differentiable torch ops on the fp32 reference-layout parameters)
# ----------------------------------------------------------------------------------------------
def pack_conv3x3(w: Tensor, cin_pad: Optional[int] = None, cout_pad: Optional[int] = None) -> Tensor:
    """OIHW [O, I, 3, 3] -> [O, 9*I] with K index (dy*3+dx)*I + i; optional zero padding of I / O."""
    o, i = w.shape[0], w.shape[1]
    t = w.permute(0, 2, 3, 1)
    if cin_pad is not None and cin_pad > i:
        t = torch.nn.functional.pad(t, (0, cin_pad - i))
        i = cin_pad
    t = t.reshape(o, 9 * i)
    if cout_pad is not None and cout_pad > o:
        t = torch.nn.functional.pad(t, (0, 0, 0, cout_pad - o))
    return t


def pack_conv1x1(w: Tensor) -> Tensor:
    return w.reshape(w.shape[0], w.shape[1])


def pack_downsample(w_s2: Tensor, w_dc: Optional[Tensor]) -> Tensor:
    """[O, 13*C]: 9 conv taps then the 4 pixel_unshuffle slabs.  pixel_unshuffle channel order is c*4+i*2+j
    (SURVEY 8 a14); the phase view delivers (i, j, c), so dc_conv's input channels are permuted here.
    ``w_dc=None`` (no DC path): [O, 9*C]."""
    if w_dc is None:
        return pack_conv3x3(w_s2)
    o, c = w_s2.shape[0], w_s2.shape[1]
    dc = w_dc.reshape(o, c, 2, 2).permute(0, 2, 3, 1).reshape(o, 4 * c)
    return torch.cat([pack_conv3x3(w_s2), dc], dim=1)


def pack_upsample_conv1(w: Tensor) -> Tensor:
    """[O, 16*I]: per output phase (py, px) the four 2x2 taps with the coinciding 3x3 taps summed (fp32)."""
    slabs = []
    for py in range(2):
        for px in range(2):
            for _, dys in _UP_ROWS[py]:
                for _, dxs in _UP_ROWS[px]:
                    s = None
                    for dy in dys:
                        for dx in dxs:
                            s = w[:, :, dy, dx] if s is None else s + w[:, :, dy, dx]
                    slabs.append(s)
    return torch.cat(slabs, dim=1)


def pack_upsample_conv2(w: Tensor, w_dc: Optional[Tensor]) -> Tensor:
    """[O, 9*Cmid (+ 4*Cin)]: conv taps then, per output phase (i, j), rows c*4+i*2+j of dc_conv."""
    parts = [pack_conv3x3(w)]
    if w_dc is not None:
        o = w.shape[0]
        dc = w_dc.reshape(o, 2, 2, w_dc.shape[1])
        for i in range(2):
            for j in range(2):
                parts.append(dc[:, i, j, :])
    return torch.cat(parts, dim=1)


def bias_upsample_conv2(b: Tensor, b_dc: Optional[Tensor]) -> Tensor:
    """[4, O] bias per output phase: conv bias + dc bias rows c*4+i*2+j."""
    if b_dc is None:
        return b.unsqueeze(0).expand(4, -1).contiguous()
    o = b.shape[0]
    dc = b_dc.reshape(o, 2, 2)
    return torch.stack([b + dc[:, i, j] for i in range(2) for j in range(2)], dim=0)


def fold_qkv(wq: Tensor, wk: Tensor, wv: Tensor, gq: Tensor, bq: Tensor, gk: Tensor, bk: Tensor, gv: Tensor,
             bv: Tensor, w1: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Fold RMSNorm weight w1 (blocks.py:146) and the three pre-projection LayerNorm affines
    (attention.py:71-73) into one [3C, C] projection:

        q = to_q(LN_q(h)),  h = x*w1/rms,  LN(h) = g*(h-mu)/sigma + b
          = a_row * (Wq*g*w1) x  -  b_row * sum_c(Wq*g)  +  Wq b        with a_row = 1/(sigma*rms), b_row = mu/sigma

    Returns (W'' [3C, C], col_sum [3C], bias' [3C]), all fp32 (the caller casts W'' to bf16).
    """
    wg = torch.cat([wq * gq, wk * gk, wv * gv], dim=0)
    bias = torch.cat([wq @ bq, wk @ bk, wv @ bv], dim=0)
    return wg * w1, wg.sum(dim=1), bias


def fold_qkv_affine(wq: Tensor, wk: Tensor, wv: Tensor, gq: Tensor, bq: Tensor, gk: Tensor, bk: Tensor, gv: Tensor,
                    bv: Tensor) -> Tuple[Tensor, Tensor]:
    """Training path: only the LayerNorm affines are folded (the normalised tensor is materialised):
    ([Wq g_q; Wk g_k; Wv g_v] [3C, C], [Wq b_q; Wk b_k; Wv b_v] [3C])."""
    return torch.cat([wq * gq, wk * gk, wv * gv], dim=0), torch.cat([wq @ bq, wk @ bk, wv @ bv], dim=0)


def rope_table(H: int, W: int, inv_freq: Tensor) -> Tensor:
    """[max(H, W), 16, 2] fp32 (cos, sin) of pos * inv_freq[k] -- attention.py:149-174 builds the same angles
    (positions in fp32, outer product with inv_freq, then cos / sin)."""
    pos = torch.arange(max(H, W), device=inv_freq.device, dtype=torch.float32)
    ang = torch.outer(pos, inv_freq.float())
    return torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()


# ----------------------------------------------------------------------------------------------
# input-gradient ("dgrad") plans: the same kernel with the taps transposed.  A = dZ (gradient w.r.t. the forward
# pre-activation output, N channels), output = dX (C channels), weights = per-tap W^T packed as [C, taps*N].
# ----------------------------------------------------------------------------------------------
def plan_conv3x3_dgrad(n: int) -> Plan:
    """dX[p] = sum_taps dZ[p - off] W_tap^T for a 3x3 stride-1 conv with ``n`` output channels."""
    taps = [TapSpec(0, 0, -(dx - 1), 0, -(dy - 1), _kb(n), (dy * 3 + dx) * n) for dy in range(3) for dx in range(3)]
    return Plan([taps], [0], [0], k_total=9 * n, name="conv3x3_dgrad", algo_k=9 * n)


def pack_conv3x3_dgrad(w: Tensor, cin_pad: Optional[int] = None, cout_pad: Optional[int] = None) -> Tensor:
    """OIHW [O, I, 3, 3] -> [I(pad), 9*O(pad)] with K index (dy*3+dx)*O + o."""
    o, i = w.shape[0], w.shape[1]
    t = w.permute(1, 2, 3, 0)              # [I, 3, 3, O]
    if cout_pad is not None and cout_pad > o:
        t = torch.nn.functional.pad(t, (0, cout_pad - o))
        o = cout_pad
    t = t.reshape(i, 9 * o)
    if cin_pad is not None and cin_pad > i:
        t = torch.nn.functional.pad(t, (0, 0, 0, cin_pad - i))
    return t


def plan_downsample_dgrad_main(c: int, n: int) -> Plan:
    """Gradient w.r.t. the stride-2 conv's input: four output phases of the [B, H, W, c] tensor, each fed by the taps
    of the forward conv that read that phase.  A = dZ [B, H/2, W/2, n] (plain)."""
    phases, out_p, out_c = [], [], []
    for p in range(2):
        for q in range(2):
            taps = []
            for dy in range(3):
                pp, dh = _s2(dy)
                if pp != p:
                    continue
                for dx in range(3):
                    qq, dw = _s2(dx)
                    if qq != q:
                        continue
                    taps.append(TapSpec(0, 0, -dw, 0, -dh, _kb(n), (dy * 3 + dx) * n))
            phases.append(taps)
            out_p.append(p)
            out_c.append(q * c)
    return Plan(phases, out_p, out_c, out_split=True, k_total=9 * n, name="downsample_dgrad_main", algo_k=9 * n // 4)


def plan_downsample_dgrad_dc(c: int, n: int) -> Plan:
    """Gradient w.r.t. x through pixel_unshuffle + 1x1 conv: phase (i, j) of dX = dZ * Wdc[:, c*4+i*2+j]."""
    phases = [[TapSpec(0, 0, 0, 0, 0, _kb(n), (i * 2 + j) * n)] for i in range(2) for j in range(2)]
    return Plan(phases, [0, 0, 1, 1], [0, c, 0, c], out_split=True, k_total=4 * n, name="downsample_dgrad_dc", algo_k=n)


def pack_downsample_dgrad_dc(w_dc: Tensor) -> Tensor:
    """dc_conv weight [N, 4C, 1, 1] (input channel c*4+i*2+j) -> [C, 4N] with K index (i*2+j)*N + n."""
    n, c4 = w_dc.shape[0], w_dc.shape[1]
    return w_dc.reshape(n, c4 // 4, 2, 2).permute(1, 2, 3, 0).reshape(c4 // 4, 4 * n)


def plan_upsample_conv1_dgrad(cin: int, cout: int) -> Plan:
    """Gradient w.r.t. the low-res input of nearest2x+conv3x3: 16 taps over the phase view of dZ [B, 2H, 2W, cout]."""
    taps = []
    wk = 0
    for py in range(2):
        for px in range(2):
            for dh, _ in _UP_ROWS[py]:
                for dw, _ in _UP_ROWS[px]:
                    taps.append(TapSpec(0, px * cout, -dw, py, -dh, _kb(cout), wk))
                    wk += cout
    return Plan([taps], [0], [0], a0_split=True, k_total=16 * cout, name="upsample_conv1_dgrad", algo_k=9 * cout * 4)


def pack_upsample_conv1_dgrad(w: Tensor) -> Tensor:
    """[Cin, 16*Co] with K index (phase*4 + tap)*Co + o, from the summed phase taps of pack_upsample_conv1."""
    co, cin = w.shape[0], w.shape[1]
    return pack_upsample_conv1(w).reshape(co, 16, cin).permute(2, 1, 0).reshape(cin, 16 * co)


def plan_upsample_dc_dgrad(cout: int) -> Plan:
    """Gradient w.r.t. x through conv1x1 + pixel_shuffle: dX = sum over the four phases of dZ_phase * Wdc_phase^T."""
    taps = [TapSpec(0, j * cout, 0, i, 0, _kb(cout), (i * 2 + j) * cout) for i in range(2) for j in range(2)]
    return Plan([taps], [0], [0], a0_split=True, k_total=4 * cout, name="upsample_dc_dgrad", algo_k=4 * cout)


def pack_upsample_dc_dgrad(w_dc: Tensor) -> Tensor:
    """dc_conv weight [4Co, Cin, 1, 1] (output channel c*4+i*2+j) -> [Cin, 4Co] with K index (i*2+j)*Co + c."""
    co4, cin = w_dc.shape[0], w_dc.shape[1]
    return w_dc.reshape(co4 // 4, 2, 2, cin).permute(3, 1, 2, 0).reshape(cin, co4)
