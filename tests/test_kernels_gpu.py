"""Per-kernel parity on a real B200: every C-ABI entry point against plain fp32 torch ops on the same inputs.

Tolerances: operands are rounded to bf16 once (inputs/weights) and accumulated in fp32 by both sides, outputs are
bf16 -> max|ours-ref|/max|ref| <= 1e-2 (2 bf16 ulps = 7.8e-3) unless stated.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from transvae import _taps as T  # noqa: E402
from transvae import ops  # noqa: E402

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def bf(x):
    return x.to(torch.bfloat16)


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("M,K,N", [(128, 64, 64), (256, 128, 128), (1000, 192, 192), (4096, 384, 1152),
                                   (300, 1536, 256), (2048, 768, 3072), (77, 64, 320)])
def test_gemm_plain(M, K, N):
    x, w, b = bf(rnd(M, K)), bf(rnd(N, K, seed=1, scale=0.05)), rnd(N, seed=2)
    y = ops.linear(x, w, T.plan_linear(K), bias=b)
    ref = x.float() @ w.float().t() + b
    assert y.shape == (M, N)
    assert rel(y, ref) < 1e-2, rel(y, ref)


def test_gemm_epilogues():
    M, K, N = 640, 256, 384
    x, w, b = bf(rnd(M, K)), bf(rnd(N, K, seed=1, scale=0.05)), rnd(N, seed=2)
    res = bf(rnd(M, N, seed=3))
    rs, rsh, cs = rnd(M, seed=4).abs() + 0.5, rnd(M, seed=5), rnd(N, seed=6)
    acc = x.float() @ w.float().t()
    y = ops.linear(x, w, T.plan_linear(K), bias=b, act=ops.ACT_GELU)
    assert rel(y, F.gelu(acc + b)) < 1e-2
    y = ops.linear(x, w, T.plan_linear(K), bias=b, act=ops.ACT_SILU)
    assert rel(y, F.silu(acc + b)) < 1e-2
    y = ops.linear(x, w, T.plan_linear(K), bias=b, residual=res)
    assert rel(y, acc + b + res.float()) < 1e-2
    y = ops.linear(x, w, T.plan_linear(K), bias=b, row_scale=rs, row_shift=rsh, col_sum=cs, residual=res)
    ref = acc * rs[:, None] - rsh[:, None] * cs[None] + b + res.float()
    assert rel(y, ref) < 1e-2
    y = ops.linear(x, w, T.plan_linear(K), bias=b, row_scale=rs, act=ops.ACT_GELU)
    assert rel(y, F.gelu(acc * rs[:, None] + b)) < 1e-2
    y = ops.linear(x, w, T.plan_linear(K))
    assert rel(y, acc) < 1e-2


@pytest.mark.parametrize("B,C,H,W,N", [(2, 64, 16, 16, 64), (1, 192, 128, 128, 192), (3, 128, 32, 32, 256),
                                       (2, 64, 8, 8, 128), (1, 64, 4, 4, 64), (2, 128, 64, 64, 128),
                                       (1, 64, 5, 256, 128), (2, 192, 3, 384, 192)])   # two / three 128-pixel tiles per row
def test_conv3x3(B, C, H, W, N):
    x, w, b = bf(rnd(B, C, H, W)), bf(rnd(N, C, 3, 3, seed=1, scale=0.05)), rnd(N, seed=2)
    res = bf(rnd(B, N, H, W, seed=3))
    wp = T.pack_conv3x3(w).contiguous()
    conv = F.conv2d(x.float(), w.float(), b, padding=1)
    y = ops.mtgemm(T.plan_conv3x3(C), nhwc(x), wp, out_shape=(B, H, W, N), bias=b, residual=nhwc(res))
    assert rel(nchw(y), conv + res.float()) < 1e-2, rel(nchw(y), conv + res.float())
    y = ops.mtgemm(T.plan_conv3x3(C), nhwc(x), wp, out_shape=(B, H, W, N), bias=b, act=ops.ACT_SILU)
    assert rel(nchw(y), F.silu(conv)) < 1e-2
    y = ops.mtgemm(T.plan_conv3x3(C), nhwc(x), wp, out_shape=(B, H, W, N), bias=b, act=ops.ACT_GELU)
    assert rel(nchw(y), F.gelu(conv)) < 1e-2


@pytest.mark.parametrize("B,C,Co,H", [(2, 64, 128, 16), (1, 192, 192, 64), (2, 128, 256, 8)])
def test_downsample(B, C, Co, H):
    x, y1 = bf(rnd(B, C, H, H)), bf(rnd(B, C, H, H, seed=9))
    w2, b2 = bf(rnd(Co, C, 3, 3, seed=1, scale=0.05)), rnd(Co, seed=2)
    wdc, bdc = bf(rnd(Co, 4 * C, 1, 1, seed=3, scale=0.05)), rnd(Co, seed=4)
    out = ops.mtgemm(T.plan_downsample(C), nhwc(y1), T.pack_downsample(w2, wdc).contiguous(), a1=nhwc(x),
                     out_shape=(B, H // 2, H // 2, Co), bias=(b2 + bdc))
    ref = F.conv2d(y1.float(), w2.float(), b2, stride=2, padding=1) + F.conv2d(F.pixel_unshuffle(x.float(), 2), wdc.float(), bdc)
    assert rel(nchw(out), ref) < 1e-2, rel(nchw(out), ref)


@pytest.mark.parametrize("B,Ci,Co,H", [(2, 128, 64, 8), (1, 192, 192, 64), (2, 256, 128, 4)])
def test_upsample(B, Ci, Co, H):
    x = bf(rnd(B, Ci, H, H))
    w1, b1 = bf(rnd(Co, Ci, 3, 3, seed=1, scale=0.05)), rnd(Co, seed=2)
    w2, b2 = bf(rnd(Co, Co, 3, 3, seed=3, scale=0.05)), rnd(Co, seed=4)
    wdc, bdc = bf(rnd(4 * Co, Ci, 1, 1, seed=5, scale=0.05)), rnd(4 * Co, seed=6)
    xn = nhwc(x)
    y = ops.mtgemm(T.plan_upsample_conv1(Ci, Co), xn, bf(T.pack_upsample_conv1(w1.float())).contiguous(),
                   out_shape=(B, 2 * H, 2 * H, Co), bias=b1[None].expand(4, -1).contiguous(), act=ops.ACT_SILU)
    ref_y = F.silu(F.conv2d(F.interpolate(x.float(), scale_factor=2.0, mode="nearest"), w1.float(), b1, padding=1))
    assert rel(nchw(y), ref_y) < 2e-2, rel(nchw(y), ref_y)   # summed taps are re-rounded to bf16
    out = ops.mtgemm(T.plan_upsample_conv2(Co, Ci), y, T.pack_upsample_conv2(w2, wdc).contiguous(), a1=xn,
                     out_shape=(B, 2 * H, 2 * H, Co), bias=T.bias_upsample_conv2(b2, bdc).contiguous())
    ref = F.conv2d(nchw(y).float(), w2.float(), b2, padding=1) + F.pixel_shuffle(F.conv2d(x.float(), wdc.float(), bdc), 2)
    assert rel(nchw(out), ref) < 1e-2, rel(nchw(out), ref)


@pytest.mark.parametrize("B,C,H,W", [(2, 128, 16, 16), (1, 192, 5, 256), (2, 64, 3, 384)])   # W >= 128: halo tiles, pair kernel
def test_direct_fp32_heads(B, C, H, W):
    x, w, b = bf(rnd(B, C, H, W)), bf(rnd(3, C, 3, 3, seed=1, scale=0.05)), rnd(3, seed=2)
    out = torch.zeros(B, 3, H, W, device=DEV)
    bias = torch.zeros(64, device=DEV)
    bias[:3] = b
    ops.mtgemm(T.plan_conv3x3(C), nhwc(x), T.pack_conv3x3(w, cout_pad=64).contiguous(), bias=bias, out_f32=out, out_n=3)
    assert rel(out, F.conv2d(x.float(), w.float(), b, padding=1)) < 2e-3


def test_qkv_rope_epilogue():
    import transvae_oracle as O
    B, C, H, W = 2, 128, 8, 4
    S = H * W
    inv = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32)).to(DEV)
    x, w = bf(rnd(B * S, C)), bf(rnd(3 * C, C, seed=1, scale=0.1))
    tab = T.rope_table(H, W, inv)
    qs = 0.125 * math.log2(math.e)
    y = ops.linear(x, w, T.plan_linear(C), rope=(tab, C, H, W, qs)).float()
    ref = (x.float() @ w.float().t()).view(B, S, 3, C // 64, 64).permute(2, 0, 3, 1, 4)   # [3, B, nh, S, 64]
    q = O.rope2d(ref[0].cpu(), H, W, inv.cpu()).to(DEV) * qs
    k = O.rope2d(ref[1].cpu(), H, W, inv.cpu()).to(DEV)
    ours = y.view(B, S, 3, C // 64, 64).permute(2, 0, 3, 1, 4)
    assert rel(ours[0], q) < 1e-2 and rel(ours[1], k) < 1e-2 and rel(ours[2], ref[2]) < 1e-2


@pytest.mark.parametrize("B,S,C", [(2, 256, 128), (1, 1024, 64), (2, 16, 128), (1, 64, 64), (1, 4096, 128), (2, 200, 64)])
def test_attention_fwd(B, S, C):
    nh = C // 64
    qkv = bf(rnd(B, S, 3 * C, scale=1.0))
    out, lse = ops.attn_fwd(qkv, B, S, C, need_lse=True)
    t = qkv.float().view(B, S, 3, nh, 64).permute(2, 0, 3, 1, 4)
    sc = (t[0] @ t[1].transpose(-1, -2)) * math.log(2.0)       # q is pre-scaled to log2 units
    ref = torch.softmax(sc, dim=-1) @ t[2]
    ref = ref.permute(0, 2, 1, 3).reshape(B, S, C)
    assert rel(out, ref) < 2e-2, rel(out, ref)
    ref_lse = torch.logsumexp(sc, dim=-1) / math.log(2.0)
    assert float((lse - ref_lse).abs().max()) < 2e-2


@pytest.mark.parametrize("B,S,C", [(24, 320, 512), (40, 128, 512), (10, 512, 1024), (100, 200, 256)],
                         ids=["odd_blocks_ragged", "single_block_items", "even_blocks", "odd_item_count"])
def test_attention_fwd_persistent_ctas_walk_several_items(B, S, C):
    """More (query-tile pair, head, image) items than SMs: every CTA of the persistent forward kernel walks two or three items.
    Odd numbers of key blocks per item flip the buffer / phase parities from item to item, one block per item takes the
    special paths (Q released by the very first S, O staging buffer reused by the next item's first block), S = 320 / 200
    leaves the second Q tile of the last pair (partly) beyond the sequence."""
    nh = C // 64
    assert ((S + 255) // 256) * nh * B > 2 * 148
    qkv = bf(rnd(B, S, 3 * C, scale=1.0))
    out, lse = ops.attn_fwd(qkv, B, S, C, need_lse=True)
    t = qkv.float().view(B, S, 3, nh, 64).permute(2, 0, 3, 1, 4)
    sc = (t[0] @ t[1].transpose(-1, -2)) * math.log(2.0)
    ref = (torch.softmax(sc, dim=-1) @ t[2]).permute(0, 2, 1, 3).reshape(B, S, C)
    assert rel(out, ref) < 2e-2, rel(out, ref)
    # per (image, head): a stale Q / K / V / O buffer of a neighbouring item would show up as one bad slice, not in the max
    err = (out.float() - ref).view(B, S, nh, 64).abs().amax(dim=(1, 3))
    scale = ref.view(B, S, nh, 64).abs().amax(dim=(1, 3)).clamp_min(1e-6)
    assert float((err / scale).max()) < 3e-2, float((err / scale).max())
    ref_lse = torch.logsumexp(sc, dim=-1) / math.log(2.0)
    assert float((lse - ref_lse).abs().max()) < 2e-2
    out2, lse2 = ops.attn_fwd(qkv, B, S, C, need_lse=True)          # and bit-reproducible
    assert torch.equal(out, out2) and torch.equal(lse, lse2)


@pytest.mark.parametrize("order", ["rising", "falling", "flat"])
def test_attention_fwd_running_max(order):
    """Scores that rise / fall steadily along the key axis: the rising case forces the lazy O rescale of the forward
    kernel at (almost) every key block, the falling case never does, the flat case has all-equal scores."""
    B, S, C = 1, 1024, 64
    g = torch.Generator(device="cpu").manual_seed(5)
    q = torch.randn(B, S, 64, generator=g) * 0.2
    k = torch.randn(B, S, 64, generator=g) * 0.2
    v = torch.randn(B, S, 64, generator=g)
    ramp = torch.linspace(0.0, 40.0, S)
    if order == "falling":
        ramp = ramp.flip(0)
    if order == "flat":
        q.zero_()
    else:
        q[..., 0] = 1.0
        k[..., 0] = ramp                    # score(i, j) ~ ramp[j] + noise (log2 units)
    qkv = bf(torch.cat([q, k, v], dim=-1).to(DEV))
    out, lse = ops.attn_fwd(qkv, B, S, C, need_lse=True)
    t = qkv.float().view(B, S, 3, 1, 64).permute(2, 0, 3, 1, 4)
    sc = (t[0] @ t[1].transpose(-1, -2)) * math.log(2.0)
    ref = (torch.softmax(sc, dim=-1) @ t[2]).permute(0, 2, 1, 3).reshape(B, S, C)
    assert rel(out, ref) < 2e-2, rel(out, ref)
    ref_lse = torch.logsumexp(sc, dim=-1) / math.log(2.0)
    assert float((lse - ref_lse).abs().max()) < 2e-2


def _gn_ref(y_nhwc, groups):
    """(sum, sum of squares) per image and channel group of the stored bf16 tensor, fp64."""
    B, H, W, C = y_nhwc.shape
    v = y_nhwc.double().reshape(B, H * W, groups, C // groups)
    return torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1)


@pytest.mark.parametrize("B,C,H,W,N,res", [
    (2, 192, 32, 32, 192, True),     # CTA-pair kernel, N = 192 (6 channels per group straddle the 8-wide vectors): fused
    (3, 128, 16, 24, 256, False),    # W not a power of two: clipped rows must stay out of the sums; odd tile count
    (2, 64, 16, 16, 128, True),      # N = 128
    (5, 64, 8, 8, 128, False),       # 64 pixels per image -> two images per tile -> library falls back to the stats pass
    (1, 64, 16, 16, 64, True),       # N = 64: 1-CTA kernel -> fallback
    (1, 320, 3, 128, 320, True),     # giant variant: 10 channels per group straddle column groups AND the 64-wide n tiles;
                                     # 128-pixel rows -> halo tiles on the N = 64 pair kernel
    (2, 64, 2, 256, 64, False),      # N = 64 with halo tiles and fused statistics (two tiles per image row)
])
def test_conv3x3_gn_sums(B, C, H, W, N, res):
    """tvae_mtgemm with gn_sums: GroupNorm(32) statistics of the convolution output from the GEMM epilogue
    (blocks.py:60,64: the norm that consumes the convolution) == statistics of the stored tensor."""
    x, w, b = bf(rnd(B, C, H, W)), bf(rnd(N, C, 3, 3, seed=1, scale=0.05)), rnd(N, seed=2)
    r = nhwc(bf(rnd(B, N, H, W, seed=3))) if res else None
    wp = T.pack_conv3x3(w).contiguous()
    y0 = ops.mtgemm(T.plan_conv3x3(C), nhwc(x), wp, out_shape=(B, H, W, N), bias=b, residual=r)
    y = ops.mtgemm(T.plan_conv3x3(C), nhwc(x), wp, out_shape=(B, H, W, N), bias=b, residual=r, gn_groups=32)
    assert torch.equal(y, y0)                          # the statistics do not disturb the output
    ref = _gn_ref(y, 32)
    got = y._gn_sums.double()
    assert got.shape == (B, 32, 2)
    err = float(((got - ref).abs() / (ref.abs() + 1.0)).max())
    assert err < 1e-4, err
    # and they drive the apply pass to the same result as the two-pass kernel
    g, be = rnd(N, seed=4) * 0.2 + 1, rnd(N, seed=5) * 0.1
    a = ops.groupnorm_silu(y, g, be, sums=y._gn_sums)
    c = ops.groupnorm_silu(y, g, be)
    assert rel(a, c) < 4e-3, rel(a, c)


def test_upsample_and_conv_in_gn_sums():
    """Phase-split output (Upsample conv2 + DC path) and the [B, 1, H*W, .] view of conv_in."""
    B, Ci, Co, H = 2, 128, 192, 16
    x, y1 = bf(rnd(B, Ci, H, H)), bf(rnd(B, Co, 2 * H, 2 * H, seed=7))
    w2, b2 = bf(rnd(Co, Co, 3, 3, seed=3, scale=0.05)), rnd(Co, seed=4)
    wdc, bdc = bf(rnd(4 * Co, Ci, 1, 1, seed=5, scale=0.05)), rnd(4 * Co, seed=6)
    out = ops.mtgemm(T.plan_upsample_conv2(Co, Ci), nhwc(y1), T.pack_upsample_conv2(w2, wdc).contiguous(), a1=nhwc(x),
                     out_shape=(B, 2 * H, 2 * H, Co), bias=T.bias_upsample_conv2(b2, bdc).contiguous(), gn_groups=32)
    ref = _gn_ref(out, 32)
    err = float(((out._gn_sums.double() - ref).abs() / (ref.abs() + 1.0)).max())
    assert err < 1e-4, err
    xi, w, b = rnd(3, 3, 16, 24), rnd(192, 3, 3, 3, seed=1, scale=0.3), rnd(192, seed=2)
    y = ops.conv_in(xi, w, b, gn_groups=32)
    assert rel(nchw(y), F.conv2d(xi, w, b, padding=1)) < 8e-3
    ref = _gn_ref(y, 32)
    err = float(((y._gn_sums.double() - ref).abs() / (ref.abs() + 1.0)).max())
    assert err < 1e-4, err


@pytest.mark.parametrize("A,B,taps", [(192, 192, 9), (384, 1536, 1), (1536, 384, 1), (64, 34, 9), (132, 68, 1), (768, 768, 9)])
def test_weight_pack_and_unpack(A, B, taps):
    """tvae_weight_pack / tvae_wgrad_unpack == the torch permutes of _taps.pack_conv3x3 / pack_conv3x3_dgrad (bit-exact:
    pure data movement plus one round-to-nearest bf16 conversion)."""
    w = rnd(A, B, 3, 3) if taps == 9 else rnd(A, B)
    wf, wd = ops.weight_pack(w, fwd=True, dgrad=True)
    if taps == 9:
        assert torch.equal(wf, bf(T.pack_conv3x3(w)).contiguous())
        assert torch.equal(wd, bf(T.pack_conv3x3_dgrad(w)).contiguous())
        g = rnd(A, 9 * B, seed=5)
        ref = g.view(A, 3, 3, B).permute(0, 3, 1, 2).contiguous()
        assert torch.equal(ops.wgrad_unpack(g, w.shape), ref)
        slot = rnd(A, B, 3, 3, seed=6)                      # accumulate into an existing gradient (a .grad slot)
        want = slot + ref
        assert ops.wgrad_unpack(g, w.shape, accumulate_into=slot) is slot and torch.equal(slot, want)
    else:
        assert torch.equal(wf, bf(w))
        assert torch.equal(wd, bf(w.t()).contiguous())
    only_f, none_d = ops.weight_pack(w, fwd=True, dgrad=False)
    assert none_d is None and torch.equal(only_f, wf)


def test_conv_in():
    x, w, b = rnd(2, 3, 32, 48), rnd(64, 3, 3, 3, seed=1, scale=0.3), rnd(64, seed=2)
    y = ops.conv_in(x, w, b)
    assert rel(nchw(y), F.conv2d(x, w, b, padding=1)) < 8e-3


@pytest.mark.parametrize("B,C,H", [(2, 192, 32), (3, 64, 16), (1, 128, 64)])
def test_groupnorm_silu(B, C, H):
    x = bf(rnd(B, C, H, H, scale=3.0) + 0.7)
    g, b = rnd(C, seed=1) * 0.2 + 1, rnd(C, seed=2) * 0.1
    y = ops.groupnorm_silu(nhwc(x), g, b)
    ref = F.silu(F.group_norm(x.float(), 32, g, b, 1e-5))
    assert rel(nchw(y), ref) < 1e-2, rel(nchw(y), ref)


def test_row_stats():
    M, C = 777, 384
    x = bf(rnd(M, C, scale=4.0) + 0.3)
    w1 = rnd(C, seed=1) * 0.2 + 1
    a, _ = ops.row_stats(x)
    xf = x.float()
    rms = torch.sqrt((xf ** 2).mean(-1) + 1e-6)
    assert rel(a, 1 / rms) < 1e-4
    a, b = ops.row_stats(x, w1)
    h = xf / rms[:, None] * w1
    mu, sigma = h.mean(-1), torch.sqrt(h.var(-1, unbiased=False) + 1e-5)
    assert rel(a, 1 / (sigma * rms)) < 1e-3 and float((b - mu / sigma).abs().max()) < 1e-3


def test_layout_reparam_loss():
    z = rnd(2, 32, 4, 4)
    zn = ops.nchw_to_nhwc(z, 64)
    assert zn.shape == (2, 4, 4, 64) and float(zn[..., 32:].abs().max()) == 0
    assert rel(ops.nhwc_to_nchw(zn, 32), bf(z).float()) == 0
    mu, lv, eps = rnd(2, 32, 4, 4, scale=40), rnd(2, 32, 4, 4, seed=1, scale=20), rnd(2, 32, 4, 4, seed=2)
    zz, mo, lo = ops.reparam(mu, lv, eps, patched=True)
    mc, lc = mu.clamp(-50, 50), lv.clamp(-30, 20)
    assert rel(zz, mc + eps * torch.exp(0.5 * lc)) < 1e-5 and torch.equal(mo, mc) and torch.equal(lo, lc)
    recon, tgt = rnd(2, 3, 16, 16, scale=3), torch.rand(2, 3, 16, 16, device=DEV)
    acc = ops.loss_sums(recon, tgt, mu, lv, patched=True)
    l1 = (recon.sigmoid() - tgt).abs().sum()
    kl = (-0.5 * (1 + lc - mu.pow(2) - lc.exp())).sum()
    assert abs(float(acc[0] - l1)) / float(l1) < 1e-4 and abs(float(acc[1] - kl)) / abs(float(kl)) < 1e-4
    assert float(acc[2]) == 0
