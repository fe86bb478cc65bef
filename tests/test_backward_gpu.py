"""Per-kernel parity of the backward pass on a real B200: every backward C-ABI entry point against torch.autograd
through the plain fp32 torch op it differentiates (inputs rounded to bf16 once; fp32 accumulation on both sides).
Tolerance: max|ours-ref|/max|ref| <= 1e-2 for bf16 outputs, 2e-3 for fp32 outputs (weight gradients)."""
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


def grads(fn, inputs, dout):
    inputs = [t.float().clone().requires_grad_(True) for t in inputs]
    fn(*inputs).backward(dout.float())
    return [t.grad for t in inputs]


@pytest.mark.parametrize("M,K,N", [(1000, 128, 192), (4096, 384, 1536), (300, 1536, 256), (8192, 64, 64), (77, 256, 320),
                                   # N and K multiples of 256: the CTA-pair weight-gradient kernel (many / few pixel tiles)
                                   (2048, 768, 768), (1000, 1536, 512), (512, 256, 256), (20000, 256, 1024)])
def test_linear_backward(M, K, N):
    x, w, dz = bf(rnd(M, K)), bf(rnd(N, K, seed=1, scale=0.05)), bf(rnd(M, N, seed=2))
    gx, gw = grads(lambda x, w: x @ w.t(), [x, w], dz)
    dx = ops.linear(dz, w.t().contiguous(), T.plan_linear(N))
    assert rel(dx, gx) < 1e-2
    dw = ops.mtgemm_wgrad(T.plan_linear(K), x.reshape(1, 1, M, K), dz.reshape(1, 1, M, N), N)
    assert rel(dw, gw) < 2e-3, rel(dw, gw)
    # bias gradient fused into the same launch (dZ^T x ones on the tensor core)
    dw2, db = ops.mtgemm_wgrad(T.plan_linear(K), x.reshape(1, 1, M, K), dz.reshape(1, 1, M, N), N, bias=True)
    assert torch.equal(dw2, dw) or rel(dw2, dw) < 1e-5
    assert db.shape == (1, N) and rel(db[0], dz.float().sum(0)) < 1e-4, rel(db[0], dz.float().sum(0))


@pytest.mark.parametrize("B,C,H,W,N", [(2, 64, 16, 16, 64), (1, 192, 128, 128, 192), (3, 128, 32, 32, 256), (2, 64, 8, 8, 128),
                                       # maps >= 128 pixels wide, N = 192 / 64: halo weight-gradient kernel -- two tiles per row,
                                       # a clipped second tile (W = 160), 1 / 2 / 3 chunks (10 / 19 / 28 slots incl. the bias slot)
                                       (2, 192, 8, 256, 192), (1, 64, 4, 160, 64), (1, 128, 2, 128, 192), (2, 64, 3, 128, 192),
                                       # C and N multiples of 256: CTA-pair kernel with nine taps
                                       (2, 256, 16, 16, 256), (1, 512, 8, 8, 768)])
def test_conv3x3_backward(B, C, H, W, N):
    x, w, dz = bf(rnd(B, C, H, W)), bf(rnd(N, C, 3, 3, seed=1, scale=0.05)), bf(rnd(B, N, H, W, seed=2))
    gx, gw = grads(lambda x, w: F.conv2d(x, w, padding=1), [x, w], dz)
    dx = ops.mtgemm(T.plan_conv3x3_dgrad(N), nhwc(dz), bf(T.pack_conv3x3_dgrad(w.float())).contiguous(), out_shape=(B, H, W, C))
    assert rel(nchw(dx), gx) < 1e-2
    dw, db = ops.mtgemm_wgrad(T.plan_conv3x3(C), nhwc(x), nhwc(dz), N, bias=True)
    assert rel(dw, T.pack_conv3x3(gw)) < 2e-3, rel(dw, T.pack_conv3x3(gw))
    assert rel(db[0], dz.float().sum(dim=(0, 2, 3))) < 1e-4


@pytest.mark.parametrize("B,C,N,H", [(2, 64, 128, 16), (1, 192, 192, 64), (1, 256, 512, 16)])     # last: CTA-pair kernel on phase views
def test_downsample_backward(B, C, N, H):
    x, y = bf(rnd(B, C, H, H)), bf(rnd(B, C, H, H, seed=5))
    w2, wdc = bf(rnd(N, C, 3, 3, seed=1, scale=0.05)), bf(rnd(N, 4 * C, 1, 1, seed=3, scale=0.05))
    dz = bf(rnd(B, N, H // 2, H // 2, seed=7))
    gy, gx, gw2, gwdc = grads(lambda y, x, w2, wdc: F.conv2d(y, w2, stride=2, padding=1) +
                              F.conv2d(F.pixel_unshuffle(x, 2), wdc), [y, x, w2, wdc], dz)
    dzn = nhwc(dz)
    dy = ops.mtgemm(T.plan_downsample_dgrad_main(C, N), dzn, bf(T.pack_conv3x3_dgrad(w2.float())).contiguous(), out_shape=(B, H, H, C))
    assert rel(nchw(dy), gy) < 1e-2
    dx = ops.mtgemm(T.plan_downsample_dgrad_dc(C, N), dzn, bf(T.pack_downsample_dgrad_dc(wdc.float())).contiguous(), out_shape=(B, H, H, C))
    assert rel(nchw(dx), gx) < 1e-2
    dw, db = ops.mtgemm_wgrad(T.plan_downsample(C), nhwc(y), dzn, N, a1=nhwc(x), bias=True)
    assert rel(dw, T.pack_downsample(gw2, gwdc)) < 2e-3
    assert rel(db[0], dz.float().sum(dim=(0, 2, 3))) < 1e-4


@pytest.mark.parametrize("B,Ci,Co,H", [(2, 128, 64, 8), (1, 192, 192, 32), (1, 512, 256, 8)])     # last: CTA-pair kernel, four phases
def test_upsample_backward(B, Ci, Co, H):
    x = bf(rnd(B, Ci, H, H))
    w1, w2 = bf(rnd(Co, Ci, 3, 3, seed=1, scale=0.05)), bf(rnd(Co, Co, 3, 3, seed=2, scale=0.05))
    wdc = bf(rnd(4 * Co, Ci, 1, 1, seed=3, scale=0.05))
    dz1 = bf(rnd(B, Co, 2 * H, 2 * H, seed=4))
    gx, gw1 = grads(lambda x, w: F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1), [x, w1], dz1)
    dx = ops.mtgemm(T.plan_upsample_conv1_dgrad(Ci, Co), nhwc(dz1), bf(T.pack_upsample_conv1_dgrad(w1.float())).contiguous(),
                    out_shape=(B, H, H, Ci))
    assert rel(nchw(dx), gx) < 2e-2     # summed taps are re-rounded to bf16
    dwp, dbp = ops.mtgemm_wgrad(T.plan_upsample_conv1(Ci, Co), nhwc(x), nhwc(dz1), Co, bias=True)
    # four output phases (p, q): db[2p + q] = sum over pixels (2h + p, 2w + q); their total is the conv bias gradient
    ph = dz1.float().view(B, Co, H, 2, H, 2).sum(dim=(0, 2, 4)).permute(1, 2, 0).reshape(4, Co)
    assert rel(dbp, ph) < 1e-4, rel(dbp, ph)
    w1r = w1.float().clone().requires_grad_(True)
    (T.pack_upsample_conv1(w1r) * dwp).sum().backward()
    assert rel(w1r.grad, gw1) < 2e-3
    y, dz2 = bf(rnd(B, Co, 2 * H, 2 * H, seed=6)), bf(rnd(B, Co, 2 * H, 2 * H, seed=8))
    gy, gx2, gw2, gwdc = grads(lambda y, x, w, wdc: F.conv2d(y, w, padding=1) + F.pixel_shuffle(F.conv2d(x, wdc), 2),
                               [y, x, w2, wdc], dz2)
    dx2 = ops.mtgemm(T.plan_upsample_dc_dgrad(Co), nhwc(dz2), bf(T.pack_upsample_dc_dgrad(wdc.float())).contiguous(),
                     out_shape=(B, H, H, Ci))
    assert rel(nchw(dx2), gx2) < 1e-2
    dw2, db2 = ops.mtgemm_wgrad(T.plan_upsample_conv2(Co, Ci), nhwc(y), nhwc(dz2), Co, a1=nhwc(x), bias=True)
    assert rel(dw2, T.pack_upsample_conv2(gw2, gwdc)) < 2e-3
    ph2 = dz2.float().view(B, Co, H, 2, H, 2).sum(dim=(0, 2, 4)).permute(1, 2, 0).reshape(4, Co)
    assert rel(db2, ph2) < 1e-4, rel(db2, ph2)


@pytest.mark.parametrize("M,K,N", [(1000, 128, 192), (4096, 1536, 384), (2048, 384, 1536)])
def test_dgrad_with_activation_gradient_epilogue(M, K, N):
    """dZ = (dY W [+ R]) * act'(Z) in the epilogue of the input-gradient GEMM (act_grad 1 / 2) == the two-step reference."""
    dy, w = bf(rnd(M, K)), bf(rnd(N, K, seed=1, scale=0.05))
    z, r = bf(rnd(M, N, seed=2, scale=2.0)), bf(rnd(M, N, seed=3))
    acc = dy.float() @ w.float().t()
    for act, fn in ((ops.ACT_GELU, F.gelu), (ops.ACT_SILU, F.silu)):
        zz = z.float().clone().requires_grad_(True)
        fn(zz).sum().backward()
        got = ops.mtgemm(T.plan_linear(K), dy.reshape(1, 1, M, K), w, out_shape=(1, 1, M, N), act=act,
                         act_grad_z=z.reshape(1, 1, M, N))
        assert rel(got.reshape(M, N), acc * zz.grad) < 1e-2, (act, rel(got.reshape(M, N), acc * zz.grad))
    zz = z.float().clone().requires_grad_(True)
    F.gelu(zz).sum().backward()
    got = ops.mtgemm(T.plan_linear(K), dy.reshape(1, 1, M, K), w, out_shape=(1, 1, M, N), act=ops.ACT_GELU,
                     residual=r.reshape(1, 1, M, N), act_grad_z=z.reshape(1, 1, M, N))
    assert rel(got.reshape(M, N), (acc + r.float()) * zz.grad) < 1e-2


def test_downsample_dgrad_with_silu_gradient_on_phase_view():
    """act_grad on a phase-split output view (the stride-2 conv's input gradient): z is staged through the same view."""
    B, C, N, H = 2, 64, 128, 16
    w2 = bf(rnd(N, C, 3, 3, seed=1, scale=0.05))
    dz, z0 = bf(rnd(B, N, H // 2, H // 2, seed=7)), bf(rnd(B, C, H, H, seed=8, scale=2.0))
    y = torch.zeros(B, C, H, H, device=DEV, requires_grad=True)
    F.conv2d(y, w2.float(), stride=2, padding=1).backward(dz.float())
    zz = z0.float().clone().requires_grad_(True)
    F.silu(zz).sum().backward()
    got = ops.mtgemm(T.plan_downsample_dgrad_main(C, N), nhwc(dz), bf(T.pack_conv3x3_dgrad(w2.float())).contiguous(),
                     out_shape=(B, H, H, C), act=ops.ACT_SILU, act_grad_z=nhwc(z0))
    assert rel(nchw(got), y.grad * zz.grad) < 1e-2


def test_bias_act_bwd_and_act_fwd():
    for (M, N) in [(777, 192), (300, 6144), (512, 64)]:
        z, dy = bf(rnd(M, N, scale=2.0)), bf(rnd(M, N, seed=1))
        for act, f in ((ops.ACT_GELU, F.gelu), (ops.ACT_SILU, F.silu)):
            zf = z.float().clone().requires_grad_(True)
            f(zf).backward(dy.float())
            dz, cs = ops.bias_act_bwd(dy, z, act)
            assert rel(dz, zf.grad) < 1e-2
            assert rel(cs, dz.float().sum(0)) < 1e-4
            assert rel(ops.act_fwd(z, act), f(z.float())) < 1e-2
        dz, cs = ops.bias_act_bwd(dy, None, ops.ACT_NONE)
        assert dz is dy and rel(cs, dy.float().sum(0)) < 1e-4
    dy = bf(rnd(2, 8, 12, 64))                       # phase view sums: [p, q, c]
    _, cs = ops.bias_act_bwd(dy, None, ops.ACT_NONE, phase_view=True)
    ref = dy.float().view(2, 4, 2, 6, 2, 64).sum(dim=(0, 1, 3))
    assert rel(cs, ref) < 1e-4


@pytest.mark.parametrize("B,C,H", [(2, 192, 32), (3, 64, 16)])
def test_groupnorm_backward(B, C, H):
    x, dh = bf(rnd(B, C, H, H, scale=3.0) + 0.7), bf(rnd(B, C, H, H, seed=3))
    g, b = rnd(C, seed=1) * 0.2 + 1, rnd(C, seed=2) * 0.1
    add = bf(rnd(B, C, H, H, seed=4))
    gx, gg, gb = grads(lambda x, g, b: F.silu(F.group_norm(x, 32, g, b, 1e-5)), [x, g, b], dh)
    xn = nhwc(x)
    sums = ops.groupnorm_stats(xn)
    dx, dg, db = ops.groupnorm_bwd(xn, nhwc(dh), sums, g, b, add=nhwc(add))
    assert rel(nchw(dx), gx + add.float()) < 1e-2
    assert rel(dg, gg) < 5e-3 and rel(db, gb) < 5e-3


@pytest.mark.parametrize("B,C,H,W", [(3, 192, 4, 128), (2, 192, 3, 256), (2, 64, 16, 16), (5, 256, 2, 128), (2, 128, 3, 200)])
def test_groupnorm_backward_reduce_from_dgrad_epilogue(B, C, H, W):
    """The input-gradient GEMM of a ResBlock convolution leaves the reduce pass of the GroupNorm backward that consumes
    its output (mtgemm gn_bwd=...; fused into the CTA-pair kernel's epilogue when a 128-pixel tile stays inside one image,
    the stand-alone reduce pass behind the GEMM otherwise): dh bit-identical to the plain launch, (dx, dgamma, dbeta)
    equal to the two-pass route and to torch (blocks.py:60-66)."""
    from transvae import _taps as T
    x, dout = bf(rnd(B, C, H, W, scale=3.0) + 0.7), bf(rnd(B, C, H, W, seed=3))
    g, b = rnd(C, seed=1) * 0.2 + 1, rnd(C, seed=2) * 0.1
    w = bf(rnd(C, C, 3, 3, seed=5) * 0.05)
    add = bf(rnd(B, C, H, W, seed=4))
    xn, dn = nhwc(x), nhwc(dout)
    sums = ops.groupnorm_stats(xn)
    wd = w.float().permute(1, 2, 3, 0).reshape(C, 9 * C).to(torch.bfloat16).contiguous()     # [Cin, tap, Cout]
    plan = T.plan_conv3x3_dgrad(C)
    dh_plain = ops.mtgemm(plan, dn, wd, out_shape=(B, H, W, C))
    dh = ops.mtgemm(plan, dn, wd, out_shape=(B, H, W, C), gn_bwd=(xn, sums, g, b, 32, 1e-5, True))
    assert torch.equal(dh, dh_plain)
    dx0, dg0, db0 = ops.groupnorm_bwd(xn, dh_plain, sums, g, b, add=nhwc(add))
    dx1, dg1, db1 = ops.groupnorm_bwd(xn, dh, sums, g, b, add=nhwc(add), part=dh._gnb_part)
    assert rel(dg1, dg0) < 1e-5 and rel(db1, db0) < 1e-5 and rel(dx1, dx0) < 2e-3
    # twice the same launch: bit-identical sums (fixed-order reduction)
    dh2 = ops.mtgemm(plan, dn, wd, out_shape=(B, H, W, C), gn_bwd=(xn, sums, g, b, 32, 1e-5, True))
    assert torch.equal(dh2._gnb_part, dh._gnb_part)
    gx, gg, gb = grads(lambda x, g, b: F.silu(F.group_norm(x, 32, g, b, 1e-5)), [x, g, b], nchw(dh))
    assert rel(nchw(dx1), gx + add.float()) < 1e-2
    assert rel(dg1, gg) < 5e-3 and rel(db1, gb) < 5e-3


def test_standalone_rmsnorm_module_forward_and_backward():
    """transvae.modules.blocks.RMSNorm called on its own (reference signature, blocks.py:168-204): 3-D [B, N, C] and 4-D
    NCHW inputs, forward and gradients against the reference formula -- one kernel each way, no eager arithmetic."""
    from transvae.modules.blocks import RMSNorm
    C = 192
    m = RMSNorm(C).to(DEV)
    with torch.no_grad():
        m.weight.copy_(rnd(C, seed=1) * 0.2 + 1)

    def ref(x, w, dim):
        return x / torch.sqrt((x ** 2).mean(dim, keepdim=True) + 1e-6) * w

    for shape, dim, wview in (((3, 50, C), -1, (C,)), ((2, C, 6, 10), 1, (1, C, 1, 1))):
        x = rnd(*shape, scale=3.0) + 0.4
        dy = rnd(*shape, seed=2)
        with torch.no_grad():
            assert rel(m(x), ref(x, m.weight.view(wview), dim)) < 1e-2
        xa = x.clone().requires_grad_(True)
        m.weight.grad = None
        m(xa).backward(dy)
        xr, wr = x.clone().requires_grad_(True), m.weight.detach().clone().requires_grad_(True)
        ref(xr, wr.view(wview), dim).backward(dy)
        assert rel(xa.grad, xr.grad) < 2e-2, (shape, rel(xa.grad, xr.grad))
        assert rel(m.weight.grad, wr.grad) < 1e-2, (shape, rel(m.weight.grad, wr.grad))


@pytest.mark.parametrize("M,C", [(500, 384), (64, 1536), (1000, 64)])
def test_token_norms(M, C):
    x, dy = bf(rnd(M, C, scale=4.0) + 0.3), bf(rnd(M, C, seed=2))
    w = rnd(C, seed=1) * 0.2 + 1
    add = bf(rnd(M, C, seed=3))

    def rms(x, w):
        return x / torch.sqrt((x ** 2).mean(-1, keepdim=True) + 1e-6) * w

    def rmsln(x, w):
        return F.layer_norm(rms(x, w), (C,), None, None, 1e-5)

    for mode, f in ((0, rms), (1, rmsln)):
        y = ops.token_norm_fwd(x, w, mode)
        assert rel(y, f(x.float(), w)) < 1e-2
        gx, gw = grads(f, [x, w], dy)
        dx, dw = ops.token_norm_bwd(x, w, dy, add, mode)
        assert rel(dx, gx + add.float()) < 1e-2, (mode, rel(dx, gx + add.float()))
        assert rel(dw, gw) < 5e-3, (mode, rel(dw, gw))


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 128), (1, 32, 32, 64), (2, 4, 4, 128), (1, 10, 20, 64)])
def test_attention_backward(B, H, W, C):
    """d(pre-RoPE q|k|v) of softmax(rope(q) rope(k)^T / 8) v against autograd through the oracle's RoPE + SDPA."""
    import transvae_oracle as O
    S, nh = H * W, C // 64
    inv = (1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32))).to(DEV)
    tab = T.rope_table(H, W, inv)
    qs = 0.125 * math.log2(math.e)
    raw = bf(rnd(B, S, 3 * C, scale=1.0))            # pre-RoPE projection output
    dout = bf(rnd(B, S, C, seed=1))
    # forward through our kernels: identity "projection" with the rope epilogue, then attention
    eye = bf(torch.eye(3 * C, device=DEV))
    qkv = ops.linear(raw.reshape(B * S, 3 * C), eye, T.plan_linear(3 * C), rope=(tab, C, H, W, qs)).reshape(B, S, 3 * C)
    out, lse = ops.attn_fwd(qkv, B, S, C, need_lse=True)
    dqkv = ops.attn_bwd(qkv, out, dout, lse, tab, B, S, C, H, W, 0.125)

    def ref_fn(r):
        t = r.view(B, S, 3, nh, 64).permute(2, 0, 3, 1, 4)
        q = O.rope2d(t[0], H, W, inv)
        k = O.rope2d(t[1], H, W, inv)
        o = F.scaled_dot_product_attention(q, k, t[2], scale=0.125)
        return o.permute(0, 2, 1, 3).reshape(B, S, C)

    r = raw.float().clone().requires_grad_(True)
    o_ref = ref_fn(r)
    o_ref.backward(dout.float())
    assert rel(out, o_ref) < 2e-2
    g = r.grad.view(B, S, 3, C)
    ours = dqkv.float().view(B, S, 3, C)
    errs = [rel(ours[:, :, i], g[:, :, i]) for i in range(3)]
    assert max(errs) < 3e-2, errs


def test_conv_in_wgrad():
    x, dy = rnd(2, 3, 32, 48), bf(rnd(2, 64, 32, 48, seed=1))
    w = rnd(64, 3, 3, 3, seed=2)
    b = rnd(64, seed=3)
    _, gw, gb = grads(lambda x, w, b: F.conv2d(x, w, b, padding=1), [x, w, b], dy)
    dw, db = ops.conv_in_wgrad(x, nhwc(dy))
    assert rel(dw, gw) < 2e-3 and rel(db, gb) < 2e-3


def test_loss_and_latent_backward():
    recon, tgt = rnd(2, 3, 16, 16, scale=3), torch.rand(2, 3, 16, 16, device=DEV)
    mu, lv, eps = rnd(2, 32, 4, 4, scale=30), rnd(2, 32, 4, 4, seed=1, scale=15), rnd(2, 32, 4, 4, seed=2)
    dz = rnd(2, 32, 4, 4, seed=3)
    r, m, l = [t.clone().requires_grad_(True) for t in (recon, mu, lv)]
    mc, lc = m.clamp(-50, 50), l.clamp(-30, 20)
    z = mc + eps * torch.exp(0.5 * lc.clamp(-30, 20))
    loss = (r.sigmoid() - tgt).abs().mean() + 1e-3 * (-0.5 * (1 + lc.clamp(-30, 20) - mc.pow(2) - lc.clamp(-30, 20).exp())).mean()
    (loss * 2.0 + (z * dz).sum()).backward()
    scal = torch.tensor([2.0 / recon.numel(), 2.0 * 1e-3 / mu.numel()], device=DEV)
    dr, dmu_c, dlv_c = ops.loss_bwd(recon, tgt, mu.clamp(-50, 50), lv.clamp(-30, 20), scal, True, (-30.0, 20.0))
    assert rel(dr, r.grad) < 1e-4
    dmu, dlv = ops.latent_bwd(mu, lv, eps, dz, dmu_c, dlv_c, True)
    assert rel(dmu, m.grad) < 1e-4 and rel(dlv, l.grad) < 1e-4


def test_adamw_step_matches_torch_adamw_with_clipping_and_warmup():
    """tvae_grad_sumsq + tvae_adamw_step == clip_grad_norm_(1.0) + torch.optim.AdamW + LambdaLR warm-up
    (train.py:608-620, train_2.py:266-274), fp32 and bf16 gradient buffers; the update counter lives on the device."""
    n = 4096 * 3
    for gdt in (torch.float32, torch.bfloat16):
        p, g = rnd(n), rnd(n, seed=1)
        ref = torch.nn.Parameter(p.clone())
        opt = torch.optim.AdamW([ref], lr=1e-2, betas=(0.9, 0.95), weight_decay=0.01)
        sch = torch.optim.lr_scheduler.LambdaLR(opt, lambda k: min(1.0, k / 2))
        m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
        state = torch.zeros(8, device=DEV)
        partials = torch.zeros(ops.SUMSQ_BLOCKS, dtype=torch.float64, device=DEV)
        ours = p.clone()
        for step in range(1, 5):
            gg = (g * step).to(gdt)
            ref.grad = gg.float().clone()
            torch.nn.utils.clip_grad_norm_([ref], 1.0)
            opt.step()
            sch.step()
            ops.grad_sumsq(gg, partials)
            ops.adamw_step(ours, gg, m, v, partials, state, 1e-2, 2, (0.9, 0.95), 1e-8, 0.01, 1.0, 1.0)
            assert abs(float(state[0]) - float(gg.float().pow(2).sum())) < 1e-3 * float(state[0])
        assert float(state[4]) == 4.0 and float(state[5]) == 0.0 and abs(float(state[6]) - 1e-2) < 1e-9
        assert rel(ours, ref.data) < 1e-4, gdt


def test_adamw_step_skips_non_finite_and_keeps_counters():
    """A non-finite gradient norm skips the update on the device: parameters, moments, the update index (bias
    correction / warm-up) all stay; only the skip counter moves (train_2.py:329-338: `continue` before optimizer.step())."""
    n = 2048
    p, g = rnd(n), rnd(n, seed=1)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    state = torch.zeros(8, device=DEV)
    partials = torch.zeros(ops.SUMSQ_BLOCKS, dtype=torch.float64, device=DEV)
    ops.grad_sumsq(g, partials)
    ops.adamw_step(p, g, m, v, partials, state, 1e-2, 0, (0.9, 0.95), 1e-8, 0.0, 1.0, 1.0)
    p1, m1 = p.clone(), m.clone()
    bad = g.clone()
    bad[7] = float("inf")
    ops.grad_sumsq(bad, partials)
    ops.adamw_step(p, bad, m, v, partials, state, 1e-2, 0, (0.9, 0.95), 1e-8, 0.0, 1.0, 1.0)
    assert torch.equal(p, p1) and torch.equal(m, m1)
    assert float(state[4]) == 1.0 and float(state[5]) == 1.0
    # the norm is formed without atomics: bit-identical on repeat
    a = partials.clone()
    ops.grad_sumsq(g, partials)
    b = partials.clone()
    ops.grad_sumsq(g, partials)
    assert torch.equal(b, partials) and not torch.equal(a, b)


def test_multi_tensor_add_and_cast():
    ts = [rnd(k, seed=k) for k in (1, 7, 192, 4096, 300000)]
    ds = [rnd(t.numel(), seed=100 + i) for i, t in enumerate(ts)]
    want = [d + t for d, t in zip(ds, ts)]
    mat = rnd(192 * 2, seed=5).view(192, 2)            # strided source: one column of a [C, 2] matrix
    d2 = rnd(192, seed=6)
    want2 = d2 + mat[:, 1]
    ops.multi_tensor_add(ds + [d2], ts + [mat[:, 1]])
    for d, w in zip(ds + [d2], want + [want2]):
        assert torch.equal(d, w)
    many = [torch.zeros(5, device=DEV) for _ in range(250)]
    ops.multi_tensor_add(many, [torch.full((5,), float(i), device=DEV) for i in range(250)])
    assert all(float(t[0]) == float(i) for i, t in enumerate(many))
    x = rnd(4096)
    y = torch.empty(4096, dtype=torch.bfloat16, device=DEV)
    ops.cast_f32_bf16(x, y)
    assert torch.equal(y, x.to(torch.bfloat16))


@pytest.mark.parametrize("C", [64, 384, 1280])
def test_fold_qkv_kernels_match_the_torch_expression(C):
    """ops.fold_qkv / fold_qkv_bwd (one launch each way) against _taps.fold_qkv_affine and its autograd gradients."""
    from transvae._autograd import FoldQkvFn
    ts = [rnd(C, C, seed=s, scale=0.05) for s in range(3)] + [rnd(C, seed=10 + s) for s in range(6)]
    a = [t.clone().requires_grad_(True) for t in ts]
    b = [t.clone().requires_grad_(True) for t in ts]
    wg, bg = FoldQkvFn.apply(*a)
    wr, br = T.fold_qkv_affine(*b)
    assert torch.equal(wg, wr) and rel(bg, br) < 1e-5
    dwg, dbg = rnd(3 * C, C, seed=20), rnd(3 * C, seed=21)
    torch.autograd.backward([wg, bg], [dwg, dbg])
    torch.autograd.backward([wr, br], [dwg, dbg])
    for i, (x, y) in enumerate(zip(a, b)):
        assert rel(x.grad, y.grad) < 2e-5, (i, rel(x.grad, y.grad))
    # fixed-order reductions: a second run is bit-identical
    a2 = [t.clone().requires_grad_(True) for t in ts]
    torch.autograd.backward(list(FoldQkvFn.apply(*a2)), [dwg, dbg])
    assert all(torch.equal(x.grad, y.grad) for x, y in zip(a, a2))
    # accumulation straight into existing gradient slots (the trainer's flat-buffer route): slot + gradient
    base = [rnd(*t.shape, seed=30 + i) for i, t in enumerate(ts)]
    slots = [t.clone() for t in base]
    ops.fold_qkv_bwd(ts[0:3], ts[3::2], ts[4::2], dwg, dbg, accumulate_into=(slots[0:3], slots[3::2], slots[4::2]))
    for i in range(9):
        assert torch.equal(slots[i], base[i] + a[i].grad), i


@pytest.mark.parametrize("O,I", [(64, 128), (192, 192)])
def test_upconv1_pack_kernels_match_the_torch_expression(O, I):
    from transvae._autograd import UpConv1PackFn
    w = rnd(O, I, 3, 3, seed=3)
    a, b = w.clone().requires_grad_(True), w.clone().requires_grad_(True)
    pa, pb = UpConv1PackFn.apply(a), T.pack_upsample_conv1(b)
    assert pa.shape == pb.shape and rel(pa, pb) < 1e-6
    d = rnd(O, 16 * I, seed=4)
    pa.backward(d)
    pb.backward(d)
    assert rel(a.grad, b.grad) < 1e-6
