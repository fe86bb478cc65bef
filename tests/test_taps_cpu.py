"""Host logic: tap tables + weight packing reproduce the reference ops (CPU, fp32, via the addressing emulator)."""
import torch
import torch.nn.functional as F

import transvae_oracle as O
from emu import emulate
from transvae import _taps as T


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def close(a, b, tol=2e-5):
    err = float((a - b).abs().max() / b.abs().max())
    assert err < tol, err


def test_linear_plan():
    x, w = rnd(3, 5, 7, 128), rnd(192, 128, seed=1)
    out = emulate(T.plan_linear(128), x.reshape(1, 1, -1, 128), None, w, (1, 1, 105, 192))
    close(out.reshape(-1, 192), x.reshape(-1, 128) @ w.t())


def test_conv3x3_plan():
    x, w, b = rnd(2, 64, 6, 10), rnd(128, 64, 3, 3, seed=1), rnd(128, seed=2)
    out = emulate(T.plan_conv3x3(64), nhwc(x), None, T.pack_conv3x3(w), (2, 6, 10, 128), b[None])
    close(nchw(out), F.conv2d(x, w, b, padding=1))


def test_conv3x3_padded_cin():
    x, w = rnd(1, 32, 4, 4), rnd(64, 32, 3, 3, seed=1)
    xp = F.pad(nhwc(x), (0, 32))
    out = emulate(T.plan_conv3x3(64), xp, None, T.pack_conv3x3(w, cin_pad=64), (1, 4, 4, 64))
    close(nchw(out), F.conv2d(x, w, padding=1))


def test_downsample_plan():
    C, Co = 64, 128
    x, y = rnd(2, C, 8, 12), rnd(2, C, 8, 12, seed=5)
    w2, b2 = rnd(Co, C, 3, 3, seed=1), rnd(Co, seed=2)
    wdc, bdc = rnd(Co, 4 * C, 1, 1, seed=3), rnd(Co, seed=4)
    ref = F.conv2d(y, w2, b2, stride=2, padding=1) + F.conv2d(F.pixel_unshuffle(x, 2), wdc, bdc)
    out = emulate(T.plan_downsample(C), nhwc(y), nhwc(x), T.pack_downsample(w2, wdc), (2, 4, 6, Co), (b2 + bdc)[None])
    close(nchw(out), ref)


def test_upsample_conv1_plan():
    Ci, Co = 128, 64
    x, w, b = rnd(2, Ci, 5, 6), rnd(Co, Ci, 3, 3, seed=1), rnd(Co, seed=2)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, b, padding=1)
    out = emulate(T.plan_upsample_conv1(Ci, Co), nhwc(x), None, T.pack_upsample_conv1(w), (2, 10, 12, Co),
                  b[None].expand(4, -1))
    close(nchw(out), ref)


def test_upsample_conv2_plan():
    Ci, Cm = 128, 64
    x, y = rnd(2, Ci, 5, 6), rnd(2, Cm, 10, 12, seed=7)
    w, b = rnd(Cm, Cm, 3, 3, seed=1), rnd(Cm, seed=2)
    wdc, bdc = rnd(4 * Cm, Ci, 1, 1, seed=3), rnd(4 * Cm, seed=4)
    ref = F.conv2d(y, w, b, padding=1) + F.pixel_shuffle(F.conv2d(x, wdc, bdc), 2)
    out = emulate(T.plan_upsample_conv2(Cm, Ci), nhwc(y), nhwc(x), T.pack_upsample_conv2(w, wdc), (2, 10, 12, Cm),
                  T.bias_upsample_conv2(b, bdc))
    close(nchw(out), ref)


def test_resblock_conv2_with_conv_shortcut_plan():
    """blocks.py:40-46, :68: conv2(h) + shortcut(x) in one accumulator, 1x1 and 3x3 shortcut; and the shortcut's input
    gradient plan (transposed taps)."""
    Ci, Co = 64, 128
    h, x = rnd(2, Co, 6, 10), rnd(2, Ci, 6, 10, seed=5)
    w2, b2 = rnd(Co, Co, 3, 3, seed=1), rnd(Co, seed=2)
    for k in (1, 3):
        ws, bs = rnd(Co, Ci, k, k, seed=3), rnd(Co, seed=4)
        ref = F.conv2d(h, w2, b2, padding=1) + F.conv2d(x, ws, bs, padding=k // 2)
        plan = T.plan_resblock_conv2(Co, Ci, k)
        out = emulate(plan, nhwc(h), nhwc(x), T.pack_resblock_conv2(w2, ws), (2, 6, 10, Co), (b2 + bs)[None])
        close(nchw(out), ref)
        assert plan.k_total == 9 * Co + k * k * Ci and len(plan.phases[0]) == 9 + k * k <= 20
        # input gradient of the shortcut: dX = conv_transpose(dZ, ws)
        dz = rnd(2, Co, 6, 10, seed=6)
        xg = x.clone().requires_grad_(True)
        F.conv2d(xg, ws, None, padding=k // 2).backward(dz)
        wd = T.pack_conv3x3_dgrad(ws) if k == 3 else ws.reshape(Co, Ci).t()
        dx = emulate(T.plan_conv_kxk_dgrad(Co, k), nhwc(dz), None, wd, (2, 6, 10, Ci))
        close(nchw(dx), xg.grad)


def test_whole_upsample_and_downsample_vs_oracle():
    cfg = dict(depths=[1, 1, 1, 1, 1], base_dims=[64, 64, 64, 128, 128])
    sd = O.init_state_dict(cfg, seed=4, mode="tamed")
    x = rnd(1, 64, 8, 8)
    p = "encoder.downsamples.2."
    y = F.silu(F.conv2d(x, sd[p + "main_path.0.weight"], sd[p + "main_path.0.bias"], padding=1))
    out = emulate(T.plan_downsample(64), nhwc(y), nhwc(x),
                  T.pack_downsample(sd[p + "main_path.2.weight"], sd[p + "dc_conv.weight"]), (1, 4, 4, 128),
                  (sd[p + "main_path.2.bias"] + sd[p + "dc_conv.bias"])[None])
    close(nchw(out), O.downsample(sd, p, x))
    p = "decoder.upsamples.1."
    x = rnd(1, 128, 4, 4, seed=3)
    y = emulate(T.plan_upsample_conv1(128, 64), nhwc(x), None, T.pack_upsample_conv1(sd[p + "main_path.1.weight"]),
                (1, 8, 8, 64), sd[p + "main_path.1.bias"][None].expand(4, -1))
    y = F.silu(y)
    out = emulate(T.plan_upsample_conv2(64, 128), y, nhwc(x),
                  T.pack_upsample_conv2(sd[p + "main_path.3.weight"], sd[p + "dc_conv.weight"]), (1, 8, 8, 64),
                  T.bias_upsample_conv2(sd[p + "main_path.3.bias"], sd[p + "dc_conv.bias"]))
    close(nchw(out), O.upsample(sd, p, x))


def test_fold_qkv_matches_rmsnorm_layernorm_linear():
    C = 128
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 9, C, generator=g) * 3 + 0.5
    w1 = 1 + 0.2 * torch.randn(C, generator=g)
    ws = [torch.randn(C, C, generator=g) * 0.05 for _ in range(3)]
    gs = [1 + 0.2 * torch.randn(C, generator=g) for _ in range(3)]
    bs = [0.1 * torch.randn(C, generator=g) for _ in range(3)]
    h = x / torch.sqrt((x ** 2).mean(-1, keepdim=True) + 1e-6) * w1
    ref = torch.cat([F.linear(F.layer_norm(h, (C,), gs[i], bs[i], 1e-5), ws[i]) for i in range(3)], dim=-1)
    W, colsum, bias = T.fold_qkv(ws[0], ws[1], ws[2], gs[0], bs[0], gs[1], bs[1], gs[2], bs[2], w1)
    rms = torch.sqrt((x ** 2).mean(-1, keepdim=True) + 1e-6)
    mu = h.mean(-1, keepdim=True)
    sigma = torch.sqrt(h.var(-1, unbiased=False, keepdim=True) + 1e-5)
    ours = (x @ W.t()) / (sigma * rms) - (mu / sigma) * colsum + bias
    close(ours, ref, 1e-4)


def test_rope_table_matches_oracle_angles():
    inv = O.init_state_dict(dict(depths=[1] * 5, base_dims=[64] * 5), 0)["encoder.stages.2.0.attn.rope.inv_freq"]
    H, W = 6, 8
    tab = T.rope_table(H, W, inv)
    ang = O.rope_angles(H, W, inv)       # [N, 64]: slots 0-15 / 16-31 rows, 32-47 / 48-63 cols
    n = 3 * W + 5
    assert torch.equal(tab[3, :, 0], ang[n, 0:16].cos()) and torch.equal(tab[3, :, 1], ang[n, 16:32].sin())
    assert torch.equal(tab[5, :, 0], ang[n, 32:48].cos()) and torch.equal(tab[5, :, 1], ang[n, 48:64].sin())
