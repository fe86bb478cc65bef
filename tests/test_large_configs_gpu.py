"""Parity at the widths and sequence lengths of the BASELINE configs the mini fixtures do not reach:

* one block per stage at the LARGE widths (ResBlock / Downsample / Upsample at 192 channels, TransVAEBlock at
  C = 384 / 768 / 1536 with 6 / 12 / 24 heads, K = 13 824 convolutions) and at the GIANT widths (320 / 640 / 1280 / 2560,
  transvae.py:135-140), each fed the same seeded input as the oracle: block output AND residual branch <= 2e-2
  (BASELINE.json north_star), block gradients <= 3e-2;
* attention forward AND backward at S = 4096 with 6 heads (the large model's first Transformer stage), forward at
  S = 16 384 and S = 65 536 (configs[3]: 512^2 / 1024^2, test_rope_extrapolation.py:28-51) against fp32 SDPA o RoPE2D;
* TransVAE-large f16d32 at 512^2 end to end against the oracle (reconstruction PSNR within 0.05 dB).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import transvae  # noqa: E402
import transvae_oracle as O  # noqa: E402
from transvae import _taps as T  # noqa: E402
from transvae import ops  # noqa: E402
from transvae.modules.blocks import ResBlock, TransVAEBlock  # noqa: E402
from transvae.modules.upsample import Downsample, Upsample  # noqa: E402
from util import build_model, nchw_f32, nhwc_bf16, rel  # noqa: E402

DEV = "cuda"
FWD_TOL, GRAD_TOL, PSNR_TOL = 2e-2, 3e-2, 0.05


def _randomised(mod, seed):
    """Reference-style init is irrelevant for parity; what matters is that every affine is non-trivial."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in mod.named_parameters():
            if p.dim() == 1 and "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / math.sqrt(fan_in)))
    return mod


def _sd(mod):
    return {k: v.detach().float().cpu().clone() for k, v in mod.state_dict().items()}


def _oracle_grads(fn, sd, x, dout):
    keys = [k for k in sd if "inv_freq" not in k]
    sdg = dict(sd)
    for k in keys:
        sdg[k] = sd[k].clone().requires_grad_(True)
    xg = x.clone().requires_grad_(True)
    out = fn(sdg, xg)
    out.backward(dout)
    return out.detach(), xg.grad, {k: sdg[k].grad for k in keys}


BLOCKS = [
    # (id, kind, channels in, channels out, B, H, W)
    ("large_res192_w128", "res", 192, 192, 1, 16, 128),       # 128-pixel rows: halo tiles, transposed N = 192 wgrad
    ("large_down192_384", "down", 192, 384, 1, 32, 32),
    ("large_up384_192", "up", 384, 192, 1, 16, 16),
    ("large_tvb384", "tvb", 384, 384, 1, 32, 32),             # 6 heads, S = 1024
    ("large_tvb768", "tvb", 768, 768, 2, 16, 16),             # 12 heads
    ("large_tvb1536", "tvb", 1536, 1536, 2, 16, 16),          # 24 heads, K = 13 824 convolution, 6144-wide hidden
    ("large_down768_1536", "down", 768, 1536, 1, 16, 16),
    ("large_up1536_768", "up", 1536, 768, 1, 8, 8),
    ("giant_res320", "res", 320, 320, 1, 8, 128),
    ("giant_down320_640", "down", 320, 640, 1, 16, 16),
    ("giant_tvb640", "tvb", 640, 640, 1, 16, 16),             # 10 heads
    ("giant_tvb1280", "tvb", 1280, 1280, 1, 16, 16),          # 20 heads
    ("giant_tvb2560", "tvb", 2560, 2560, 1, 8, 8),            # 40 heads, K = 23 040
    ("giant_up2560_1280", "up", 2560, 1280, 1, 4, 4),
]


@pytest.mark.parametrize("name,kind,cin,cout,B,H,W", BLOCKS, ids=[b[0] for b in BLOCKS])
def test_block_at_model_width_forward_and_gradients(name, kind, cin, cout, B, H, W):
    torch.manual_seed(0)
    if kind == "res":
        mod = ResBlock(cin, cout)
        fn = lambda s, t: O.resblock(s, "", t)
    elif kind == "tvb":
        mod = TransVAEBlock(cin)
        fn = lambda s, t: O.transvae_block(s, "", t, 64)
    elif kind == "down":
        mod = Downsample(cin, cout)
        fn = lambda s, t: O.downsample(s, "", t)
    else:
        mod = Upsample(cin, cout)
        fn = lambda s, t: O.upsample(s, "", t)
    mod = _randomised(mod, seed=len(name)).to(DEV)
    sd = _sd(mod)
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(B, cin, H, W, generator=g) * 2).to(torch.bfloat16).float()
    with torch.no_grad():
        out_ref0 = fn(sd, x)
    dout = torch.randn(out_ref0.shape, generator=g).to(torch.bfloat16).float()
    out_ref, dx_ref, gref = _oracle_grads(fn, sd, x, dout)
    errs = {}
    # inference path (folded norms, fused epilogues): block output and the isolated residual branch(es)
    mod.eval()
    with torch.no_grad():
        xin = nhwc_bf16(x).to(DEV)
        errs["out"] = rel(nchw_f32(mod.forward_nhwc(xin)), out_ref)
        if kind == "res":
            errs["branch"] = rel(nchw_f32(mod.forward_nhwc(xin, add_residual=False)), out_ref - x)
        elif kind == "tvb":
            a_ref = O.attention(sd, "attn.", O.rmsnorm(x, sd["norm1.weight"]), 64)
            a = mod.attn.forward_fused(xin, mod.norm1.weight, add_residual=False)
            errs["attn_branch"] = rel(nchw_f32(a), a_ref)
            x1 = x + a_ref
            f_ref = O.conv_ffn(sd, "ffn.", O.rmsnorm(x1, sd["norm2.weight"]))
            f = mod.ffn.forward_fused(nhwc_bf16(x1).to(DEV), mod.norm2.weight, add_residual=False)
            errs["ffn_branch"] = rel(nchw_f32(f), f_ref)
    # training path: forward again (materialised norms, saved activations) + hand-written backward
    mod.train()
    mod.zero_grad()
    xt = nhwc_bf16(x).to(DEV).requires_grad_(True)
    out = mod.forward_nhwc(xt)
    out.backward(nhwc_bf16(dout).to(DEV))
    errs["train_out"] = rel(nchw_f32(out.detach()), out_ref)
    errs["dx"] = rel(nchw_f32(xt.grad), dx_ref)
    gerrs = {k: rel(p.grad, gref[k]) for k, p in mod.named_parameters()}
    print(name, {k: round(v, 4) for k, v in errs.items()}, "worst grad", max(gerrs.items(), key=lambda kv: kv[1]))
    bad = {k: v for k, v in errs.items() if v > (GRAD_TOL if k == "dx" else FWD_TOL)}
    bad.update({k: v for k, v in gerrs.items() if v > GRAD_TOL})
    assert not bad, bad


VARIANTS = [
    ("convffn_depthwise_384", "ffn_dw", 384, 384, 2, 16, 16),
    ("convffn_depthwise_128_ragged", "ffn_dw", 128, 128, 1, 10, 12),
    ("resblock_shortcut1x1_192_384", "res_sc1", 192, 384, 1, 16, 128),
    ("resblock_shortcut3x3_64_128", "res_sc3", 64, 128, 2, 16, 16),
    ("resblock_shortcut1x1_128_64", "res_sc1", 128, 64, 2, 8, 8),
]


@pytest.mark.parametrize("name,kind,cin,cout,B,H,W", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_module_level_variants_forward_and_gradients(name, kind, cin, cout, B, H, W):
    """ConvFFN(conv_type='depthwise') (conv.py:42-50) and ResBlock with a 1x1 / 3x3 convolutional
    shortcut (blocks.py:40-46) -- the oracle is bit-exact to the reference for both (validate_against_reference.py)."""
    from transvae.modules.conv import ConvFFN
    torch.manual_seed(0)
    if kind == "ffn_dw":
        ffn = _randomised(ConvFFN(cin, conv_type="depthwise"), seed=3).to(DEV)
        w2 = (1.0 + 0.2 * torch.randn(cin, generator=torch.Generator().manual_seed(4))).to(DEV).requires_grad_(True)
        sd = _sd(ffn)
        sd["norm.weight"] = w2.detach().cpu().clone()
        fn = lambda s, t: t + O.conv_ffn({k: v for k, v in s.items() if k != "norm.weight"}, "", O.rmsnorm(t, s["norm.weight"]))
        fwd = lambda t: ffn.forward_fused(t, w2)
        params = dict(ffn.named_parameters())
        params["norm.weight"] = w2
    else:
        mod = _randomised(ResBlock(cin, cout, use_conv_shortcut=(kind == "res_sc3")), seed=5).to(DEV)
        sd = _sd(mod)
        fn = lambda s, t: O.resblock(s, "", t)
        fwd = mod.forward_nhwc
        params = dict(mod.named_parameters())
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(B, cin, H, W, generator=g) * 2).to(torch.bfloat16).float()
    with torch.no_grad():
        shape = fn(sd, x).shape
    dout = torch.randn(shape, generator=g).to(torch.bfloat16).float()
    out_ref, dx_ref, gref = _oracle_grads(fn, sd, x, dout)
    errs = {}
    with torch.no_grad():
        errs["out"] = rel(nchw_f32(fwd(nhwc_bf16(x).to(DEV))), out_ref)
    for p in params.values():
        p.grad = None
    xt = nhwc_bf16(x).to(DEV).requires_grad_(True)
    out = fwd(xt)
    out.backward(nhwc_bf16(dout).to(DEV))
    errs["train_out"] = rel(nchw_f32(out.detach()), out_ref)
    errs["dx"] = rel(nchw_f32(xt.grad), dx_ref)
    gerrs = {k: rel(p.grad, gref[k]) for k, p in params.items()}
    print(name, {k: round(v, 4) for k, v in errs.items()}, "worst grad", max(gerrs.items(), key=lambda kv: kv[1]))
    bad = {k: v for k, v in errs.items() if v > (GRAD_TOL if k == "dx" else FWD_TOL)}
    bad.update({k: v for k, v in gerrs.items() if v > GRAD_TOL})
    assert not bad, bad


def test_depthwise_kernels_against_conv2d():
    B, H, W, C = 2, 9, 13, 256
    g = torch.Generator().manual_seed(1)
    u = torch.randn(B, C, H, W, generator=g).to(torch.bfloat16)
    w = torch.randn(C, 1, 3, 3, generator=g) * 0.3
    b = torch.randn(C, generator=g) * 0.1
    dy = torch.randn(B, C, H, W, generator=g).to(torch.bfloat16)
    w9c = w.reshape(C, 9).t().contiguous().to(DEV)
    un, dyn = nhwc_bf16(u.float()).to(DEV), nhwc_bf16(dy.float()).to(DEV)
    y = ops.dwconv3x3(un, w9c, b.to(DEV))
    uf = u.float().requires_grad_(True)
    wf, bf_ = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = uf + F.conv2d(uf, wf, bf_, padding=1, groups=C)
    ref.backward(dy.float())
    assert rel(nchw_f32(y), ref.detach()) < 1e-2
    du = ops.dwconv3x3(dyn, w9c, None, flip=True)
    assert rel(nchw_f32(du), uf.grad) < 1e-2
    dw, db = ops.dwconv3x3_wgrad(un, dyn)
    assert rel(dw.t().reshape(C, 1, 3, 3), wf.grad) < 2e-3 and rel(db, bf_.grad) < 2e-3


def _rope_sdpa_ref(raw, B, S, C, H, W, inv):
    nh = C // 64
    t = raw.view(B, S, 3, nh, 64).permute(2, 0, 3, 1, 4)
    q = O.rope2d(t[0], H, W, inv)
    k = O.rope2d(t[1], H, W, inv)
    o = F.scaled_dot_product_attention(q, k, t[2], scale=0.125)
    return o.permute(0, 2, 1, 3).reshape(B, S, C)


def _our_attention(raw, B, S, C, H, W, inv, need_lse=True):
    tab = T.rope_table(H, W, inv)
    qs = 0.125 * math.log2(math.e)
    eye = torch.eye(3 * C, device=DEV).to(torch.bfloat16)
    qkv = ops.linear(raw.reshape(B * S, 3 * C), eye, T.plan_linear(3 * C), rope=(tab, C, H, W, qs)).reshape(B, S, 3 * C)
    out, lse = ops.attn_fwd(qkv, B, S, C, need_lse=need_lse)
    return qkv, out, lse, tab


def test_attention_fwd_bwd_s4096_six_heads():
    """The large model's first Transformer stage at 256^2: S = 4096, 6 heads (attention.py:76-92)."""
    B, H, W, C = 1, 64, 64, 384
    S = H * W
    inv = (1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32))).to(DEV)
    g = torch.Generator(device="cpu").manual_seed(3)
    raw = torch.randn(B, S, 3 * C, generator=g).to(DEV).to(torch.bfloat16)
    dout = torch.randn(B, S, C, generator=g).to(DEV).to(torch.bfloat16)
    qkv, out, lse, tab = _our_attention(raw, B, S, C, H, W, inv)
    dqkv = ops.attn_bwd(qkv, out, dout, lse, tab, B, S, C, H, W, 0.125)
    torch.backends.cuda.matmul.allow_tf32 = False
    r = raw.float().clone().requires_grad_(True)
    with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.MATH):
        o_ref = _rope_sdpa_ref(r, B, S, C, H, W, inv)
        o_ref.backward(dout.float())
    assert rel(out, o_ref) < FWD_TOL, rel(out, o_ref)
    gr = r.grad.view(B, S, 3, C)
    ours = dqkv.float().view(B, S, 3, C)
    errs = [rel(ours[:, :, i], gr[:, :, i]) for i in range(3)]
    print("attention S=4096 h=6: out", rel(out, o_ref), "dq/dk/dv", errs)
    assert max(errs) < GRAD_TOL, errs


@pytest.mark.parametrize("H,W,C", [(128, 128, 128), (256, 256, 64)], ids=["S16384_512px", "S65536_1024px"])
def test_attention_fwd_long_sequences(H, W, C):
    """configs[3]: 512^2 -> S = 16 384, 1024^2 -> S = 65 536 in the first Transformer stage.  Reference in fp32 in query
    chunks (the full S x S score matrix does not fit at 65 536)."""
    B, S, nh = 1, H * W, C // 64
    inv = (1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32))).to(DEV)
    raw = torch.randn(B, S, 3 * C, generator=torch.Generator(device="cpu").manual_seed(4)).to(DEV).to(torch.bfloat16)
    _, out, lse, _ = _our_attention(raw, B, S, C, H, W, inv)
    t = raw.float().view(B, S, 3, nh, 64).permute(2, 0, 3, 1, 4)
    q, k, v = O.rope2d(t[0], H, W, inv), O.rope2d(t[1], H, W, inv), t[2]
    ref = torch.empty(B, nh, S, 64, device=DEV)
    ref_lse = torch.empty(B, nh, S, device=DEV)
    for s0 in range(0, S, 4096):
        sc = (q[:, :, s0:s0 + 4096] @ k.transpose(-1, -2)) * 0.125
        ref[:, :, s0:s0 + 4096] = torch.softmax(sc, dim=-1) @ v
        ref_lse[:, :, s0:s0 + 4096] = torch.logsumexp(sc, dim=-1) / math.log(2.0)
    ref = ref.permute(0, 2, 1, 3).reshape(B, S, C)
    e = rel(out, ref)
    print(f"attention forward S={S}: max-rel {e:.4f}, lse max abs {float((lse - ref_lse).abs().max()):.4f}")
    assert e < FWD_TOL, e
    assert float((lse - ref_lse).abs().max()) < 3e-2


def test_large_f16d32_parity_at_512():
    """BASELINE configs[3] (resolution extrapolation, test_rope_extrapolation.py:28-51) at 512^2, batch 1: latents and
    reconstruction against the fp32 oracle; PSNR of both flavours within 0.05 dB."""
    cfg = O.variant_config("large")
    sd = O.init_state_dict(cfg, seed=0, mode="reference")
    m = build_model(cfg, sd)
    x = torch.rand(1, 3, 512, 512, generator=torch.Generator().manual_seed(77))
    with torch.no_grad():
        mu_o, lv_o = O.encode(sd, cfg, x)
        rec_o = O.decode(sd, cfg, mu_o)
        mu, lv = m.encode(x.cuda())
        rec = m.decode(mu)
        rec_dec = m.decode(mu_o.cuda())
    e = dict(mu=rel(mu, mu_o), logvar=rel(lv, lv_o), recon_e2e=rel(rec, rec_o), recon_dec=rel(rec_dec, rec_o))
    print("large@512:", e)
    assert mu.shape == (1, 32, 32, 32) and rec.shape == x.shape
    assert max(e.values()) < 4e-2, e
    for f in (lambda t: t.clamp(0, 1), torch.sigmoid):
        a, b = O.psnr(f(rec.float().cpu()), x), O.psnr(f(rec_o), x)
        print("PSNR ours / oracle", a, b)
        assert abs(a - b) < PSNR_TOL
