"""Model-level parity on a real B200 against the golden reference outputs and the oracle.

Tolerances (BASELINE.json north_star): bf16 activations within 2e-2 max relative error per block -- checked on
the block output AND on the residual branch, each block fed with the oracle's fp32 input so errors do
not accumulate; reconstruction PSNR within 0.05 dB of the fp32 reference.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

import transvae  # noqa: E402
import transvae_oracle as O  # noqa: E402
from util import build_model, load_golden, nchw_f32, nhwc_bf16, rel  # noqa: E402

BLOCK_TOL = 2e-2
PSNR_TOL = 0.05


def psnr_pair(recon, x):
    return O.psnr(recon.clamp(0, 1), x), O.psnr(recon.sigmoid(), x)


@pytest.mark.parametrize("name", ["mini_ref", "mini_tamed", "mini_tamed_128"])
def test_golden_encode_decode(name):
    blob, sd = load_golden(name)
    m = build_model(blob["cfg"], sd)
    x = blob["x"].cuda()
    with torch.no_grad():
        mu, logvar = m.encode(x)
        recon = m.decode(blob["mu"].cuda())          # decoder fed with the reference's mu: isolates the decoder
        recon_e2e = m.decode(mu)
    e = dict(mu=rel(mu, blob["mu"]), logvar=rel(logvar, blob["logvar"]), recon=rel(recon, blob["recon_from_mu"]),
             recon_e2e=rel(recon_e2e, blob["recon_from_mu"]))
    print(name, e)
    assert mu.dtype == torch.float32 and mu.shape == blob["mu"].shape and recon.shape == blob["x"].shape
    assert e["mu"] < 5e-2 and e["logvar"] < 5e-2 and e["recon"] < 5e-2 and e["recon_e2e"] < 8e-2, e
    for ours, ref in zip(psnr_pair(recon_e2e.cpu(), blob["x"]), psnr_pair(blob["recon_from_mu"], blob["x"])):
        assert abs(ours - ref) < PSNR_TOL, (ours, ref)


@pytest.mark.parametrize("name", ["mini_ref", "mini_tamed"])
def test_golden_patched_forward_and_loss(name):
    import transvae
    blob, sd = load_golden(name)
    m = build_model(blob["cfg"], sd, patched=True)
    x = blob["x"].cuda()
    with torch.no_grad():
        recon, mu, logvar = m(x, eps=blob["eps"].cuda())
        out = m(x, return_dict=True, eps=blob["eps"].cuda())
        loss = transvae.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)(
            recon, x, mu, logvar)
    assert set(out) == {"reconstruction", "mu", "logvar", "z"}
    assert float(mu.abs().max()) <= 50 and float(logvar.max()) <= 20 and float(logvar.min()) >= -30
    e = dict(recon=rel(recon, blob["recon_patched"]), mu=rel(mu, blob["mu_patched"]), logvar=rel(logvar, blob["logvar_patched"]))
    print(name, e, float(loss["total"]), float(blob["loss_total"]))
    if name == "mini_tamed":
        assert e["recon"] < 8e-2 and e["mu"] < 5e-2 and e["logvar"] < 5e-2, e
        assert abs(float(loss["l1"]) - float(blob["loss_l1"])) < 2e-3
        assert abs(float(loss["kl"]) - float(blob["loss_kl"])) <= 5e-2 * abs(float(blob["loss_kl"])) + 1e-12
    assert torch.isfinite(recon).all() and set(loss) == {"l1", "lpips", "kl", "vf", "gan", "total"}
    # the loss kernel itself, on the reference's tensors: exact formula check
    with torch.no_grad():
        l2 = transvae.TransVAELoss(lpips_weight=0.0, vf_weight=0.0, gan_weight=0.0)(
            blob["recon_patched"].cuda(), x, blob["mu_patched"].cuda(), blob["logvar_patched"].cuda())
    assert abs(float(l2["l1"]) - float(blob["loss_l1"])) < 1e-5
    assert abs(float(l2["kl"]) - float(blob["loss_kl"])) <= 1e-4 * abs(float(blob["loss_kl"]))


@pytest.mark.parametrize("name", ["mini_ref", "mini_tamed", "mini_tamed_128"])
def test_per_block_parity(name):
    """Every block of the model, fed with the ORACLE's input for that block, within 2e-2 on output and branch."""
    blob, sd = load_golden(name)
    cfg = blob["cfg"]
    m = build_model(cfg, sd)
    tr = {}
    with torch.no_grad():
        mu_o, _ = O.encode(sd, cfg, blob["x"], tr)
        O.decode(sd, cfg, mu_o, tr)
    worst = {}

    def check(key, ours_nhwc, ref):
        worst[key] = rel(nchw_f32(ours_nhwc), ref)

    with torch.no_grad():
        for side, mod in (("encoder", m.encoder), ("decoder", m.decoder)):
            prev = tr[f"{side}.conv_in"]
            n_res = (0, 1) if side == "encoder" else (len(cfg["depths"]) - 2, len(cfg["depths"]) - 1)
            for i, stage in enumerate(mod.stages):
                for j, block in enumerate(stage):
                    key = f"{side}.stages.{i}.{j}"
                    xin = nhwc_bf16(prev).cuda()
                    check(key, block.forward_nhwc(xin), tr[key])
                    # residual branches in isolation (the stream dwarfs them)
                    if i in n_res:
                        check(key + "#branch", block.forward_nhwc(xin, add_residual=False), tr[key + ".branch"])
                    else:
                        a = block.attn.forward_fused(xin, block.norm1.weight, add_residual=False)
                        check(key + "#attn_branch", a, tr[key + ".attn_branch"])
                        x1 = nhwc_bf16(prev + tr[key + ".attn_branch"]).cuda()
                        f = block.ffn.forward_fused(x1, block.norm2.weight, add_residual=False)
                        check(key + "#ffn_branch", f, tr[key + ".ffn_branch"])
                    prev = tr[key]
                samplers = mod.downsamples if side == "encoder" else mod.upsamples
                if i < len(samplers):
                    key = f"{side}.{'downsamples' if side == 'encoder' else 'upsamples'}.{i}"
                    check(key, samplers[i].forward_nhwc(nhwc_bf16(prev).cuda()), tr[key])
                    prev = tr[key]
    bad = {k: v for k, v in worst.items() if v > BLOCK_TOL}
    print(name, "worst block error:", max(worst.items(), key=lambda kv: kv[1]))
    assert not bad, bad


def test_streamed_reconstructor_matches_direct_calls():
    """transvae.streaming.StreamedReconstructor (inference_example.py:56-73 with the copies on side streams): three
    different pinned batches, pipelined with prefetch, must give what encode -> decode gives batch by batch, and an
    unannounced batch (no prefetch) must work too."""
    from transvae.streaming import StreamedReconstructor
    blob, sd = load_golden("mini_tamed")
    m = build_model(blob["cfg"], sd)
    g = torch.Generator().manual_seed(11)
    xs = [torch.rand(blob["x"].shape, generator=g).pin_memory() for _ in range(3)]
    outs = [torch.empty(blob["x"].shape, dtype=torch.float32).pin_memory() for _ in range(4)]
    pipe = StreamedReconstructor(m)
    for i, x in enumerate(xs):
        pipe.reconstruct(x, outs[i], next_x_host=xs[i + 1] if i + 1 < len(xs) else None)
    pipe.reconstruct(xs[1], outs[3])                 # not prefetched
    pipe.synchronize()
    with torch.no_grad():
        refs = [m.decode(m.encode(x.cuda())[0]).cpu() for x in xs]
        refs2 = [m.decode(m.encode(x.cuda())[0]).cpu() for x in xs]

    def l2(a, b):
        return float((a - b).norm() / b.norm())

    # The kernels are the same, but the fp32 atomics of the GroupNorm statistics reorder from run to run and a flipped
    # bf16 rounding shows up as a few-percent error at single pixels of this random-init net (max-norm), so the
    # comparison is in the l2 norm and against the measured run-to-run noise; a stream-ordering bug (input read before its
    # upload, output downloaded before the last kernel) gives errors of order 1.
    noise = max(l2(a, b) for a, b in zip(refs2, refs))
    errs = [l2(o, r) for o, r in zip(outs[:3], refs)] + [l2(outs[3], refs[1])]
    print("streamed vs direct l2", errs, "run-to-run noise", noise)
    assert max(errs) < max(2e-3, 4.0 * noise), (errs, noise)
    with pytest.raises(ValueError):
        pipe.reconstruct(torch.rand(blob["x"].shape), outs[0])      # unpinned input


def test_inference_forward_is_bit_reproducible():
    """Two runs of encode -> decode on the same input give identical bits.  The only order-dependent reduction of the
    inference path, the GroupNorm statistics, is accumulated in fp64 (fp32 tile partials in a fixed order + fp64
    atomics, E[x^2] - mean^2 in fp64); with fp32 sums identical runs differed by 1.3 % (l2) on the reconstruction of
    large-f16d32 (profiles/r1c_noise_probe_large_fp32_sums.txt)."""
    blob, sd = load_golden("mini_tamed_128")
    m = build_model(blob["cfg"], sd)
    x = blob["x"].cuda()
    with torch.no_grad():
        outs = []
        for _ in range(3):
            mu, logvar = m.encode(x)
            outs.append((mu.clone(), logvar.clone(), m.decode(mu).clone()))
    for o in outs[1:]:
        for a, b in zip(o, outs[0]):
            assert torch.equal(a, b)


def test_public_module_forward_nchw():
    """The reference's per-module public signature (NCHW in, NCHW out) on a bare block / attention / FFN."""
    blob, sd = load_golden("mini_tamed")
    cfg = blob["cfg"]
    m = build_model(cfg, sd)
    x = torch.randn(2, 64, 16, 16, generator=torch.Generator().manual_seed(3)) * 2
    with torch.no_grad():
        for key, mod, ref in [
            ("block", m.encoder.stages[2][0], O.transvae_block(sd, "encoder.stages.2.0.", x, 64)),
            ("attn", m.encoder.stages[2][0].attn, O.attention(sd, "encoder.stages.2.0.attn.", x, 64)),
            ("ffn", m.encoder.stages[2][0].ffn, O.conv_ffn(sd, "encoder.stages.2.0.ffn.", x)),
            ("res", m.encoder.stages[0][0], O.resblock(sd, "encoder.stages.0.0.", x)),
            ("down", m.encoder.downsamples[2], O.downsample(sd, "encoder.downsamples.2.", x)),
            ("rms", m.encoder.stages[2][0].norm1, O.rmsnorm(x, sd["encoder.stages.2.0.norm1.weight"])),
        ]:
            out = mod(x.cuda())
            assert out.shape == ref.shape and out.dtype == torch.float32
            assert rel(out, ref) < BLOCK_TOL, (key, rel(out, ref))


def test_resolutions_and_batch_shapes():
    """T/test_installation.py:90-113 (resolutions 128/256/512 -> latent H/16) on the tiny variant."""
    import transvae
    torch.manual_seed(0)
    m = transvae.TransVAE(variant="tiny", compression_ratio=16, latent_dim=32, input_resolution=256).cuda().eval()
    with torch.no_grad():
        for res, B in ((128, 3), (256, 2), (512, 1)):
            x = torch.rand(B, 3, res, res, device="cuda")
            mu, logvar = m.encode(x)
            assert mu.shape == (B, 32, res // 16, res // 16) == logvar.shape
            rec = m.decode(mu)
            assert rec.shape == x.shape and torch.isfinite(rec).all()


LARGE_TOL = 4e-2      # measured [B200]: mu 0.013, logvar 0.015, recon 0.022 (decoder alone 0.013)


def test_large_f16d32_parity_at_256():
    """BASELINE.json configs[1] model at full size (B=1 so the CPU oracle finishes in seconds)."""
    cfg = O.variant_config("large")
    sd = O.init_state_dict(cfg, seed=0, mode="reference")
    m = build_model(cfg, sd)
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        mu_o, lv_o = O.encode(sd, cfg, x)
        rec_o = O.decode(sd, cfg, mu_o)
        mu, lv = m.encode(x.cuda())
        rec = m.decode(mu)
        rec_dec = m.decode(mu_o.cuda())
    e = dict(mu=rel(mu, mu_o), logvar=rel(lv, lv_o), recon_e2e=rel(rec, rec_o), recon_dec=rel(rec_dec, rec_o))
    print("large@256:", e)
    # end to end through 34 blocks at bf16: the reference's own bf16-autocast forward is 2.6e-2 off its fp32 result here
    assert max(e.values()) < LARGE_TOL, e
    for ours, ref in zip(psnr_pair(rec.cpu(), x), psnr_pair(rec_o, x)):
        print("PSNR ours/ref", ours, ref)
        assert abs(ours - ref) < PSNR_TOL


@pytest.mark.parametrize("shape", [(1, 3, 96, 160), (3, 3, 128, 128), (2, 3, 80, 48), (1, 3, 32, 32)])
def test_ragged_resolutions_and_batches_vs_oracle(shape):
    """Non-square, non-power-of-two and tiny inputs: partial pixel tiles (TMA clipping / zero fill), sequence lengths
    that are not multiples of the 128-key attention block (960, 240, 60, ...), batch 1 and 3."""
    blob, sd = load_golden("mini_tamed")
    cfg = blob["cfg"]
    m = build_model(cfg, sd)
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        mu_o, lv_o = O.encode(sd, cfg, x)
        rec_o = O.decode(sd, cfg, mu_o)
        mu, lv = m.encode(x.cuda())
        rec = m.decode(mu_o.cuda())
    e = dict(mu=rel(mu, mu_o), logvar=rel(lv, lv_o), recon=rel(rec, rec_o))
    print(shape, e)
    assert mu.shape == mu_o.shape and rec.shape == rec_o.shape
    assert max(e.values()) < 5e-2, e
    # training path on the same ragged shape: loss matches the oracle's, gradients finite
    import transvae
    m.train()
    eps = torch.randn(mu_o.shape, generator=torch.Generator().manual_seed(10))
    r, mu_t, lv_t = m(x.cuda(), eps=eps.cuda())
    loss = transvae.TransVAELoss(lpips_weight=0.0, vf_weight=0.0, gan_weight=0.0)(r, x.cuda(), mu_t, lv_t)
    loss["total"].backward()
    rec_p, mu_p, lv_p, _ = O.forward(sd, cfg, x, eps, patched=True)
    ref = O.loss_l1_kl(rec_p, x, mu_p, lv_p, 1.0, 1e-8, patched=True)
    assert abs(float(loss["total"].detach()) - float(ref["total"])) < 1e-2
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())


@pytest.mark.parametrize("flags", [dict(use_rope=False), dict(use_dc_path=False), dict(use_rope=False, use_dc_path=False),
                                   dict(depths=[1, 1, 1, 2], base_dims=[64, 64, 128, 128])])   # 4-stage (f8) layout
def test_ablation_variants_forward_and_backward(flags):
    """use_rope=False / use_dc_path=False (transvae.py:36-38) and the 4-stage f8 layout
    (transvae.py:141-146) against the oracle (bit-identical to the
    reference for these flags, oracle/validate_against_reference.py): encode / decode within the bf16 block tolerance
    accumulated over the mini model, PSNR within 0.05 dB, and a training step whose gradients are finite and close."""
    cfg = dict(depths=[1, 1, 1, 1, 2], base_dims=[64, 64, 64, 128, 128], mlp_ratio=1.0, head_dim=64, latent_dim=32)
    cfg.update(flags)
    sd = O.init_state_dict(cfg, seed=2, mode="tamed")
    m = build_model(cfg, sd)
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        mu, lv = m.encode(x.cuda())
        rec = m.decode(mu)
        mu_o, lv_o = O.encode(sd, cfg, x)
        rec_o = O.decode(sd, cfg, mu_o)
        rec_same_z = O.decode(sd, cfg, mu.float().cpu())
    assert rel(mu, mu_o) < 5e-2 and rel(lv, lv_o) < 5e-2, (rel(mu, mu_o), rel(lv, lv_o))
    assert rel(rec, rec_same_z) < 5e-2, rel(rec, rec_same_z)
    for f in (lambda t: t.clamp(0, 1), torch.sigmoid):
        assert abs(O.psnr(f(rec.float().cpu()), x) - O.psnr(f(rec_o), x)) < 0.05
    # training path
    m.train()
    eps = torch.randn(mu.shape, generator=torch.Generator().manual_seed(9))
    r, mu_t, lv_t = m(x.cuda(), eps=eps.cuda())
    loss = transvae.TransVAELoss(lpips_weight=0.0, vf_weight=0.0, gan_weight=0.0)(r, x.cuda(), mu_t, lv_t)["total"]
    loss.backward()
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "inv_freq" not in k else v) for k, v in sd.items()}
    ro, muo, lvo, _ = O.forward(sdg, cfg, x, eps, patched=True)
    lo = O.loss_l1_kl(ro, x, muo, lvo, 1.0, 1e-8, patched=True)["total"]
    lo.backward()
    assert abs(float(loss) - float(lo)) < 5e-3
    for k, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
    # gradient parity, tensor by tensor, with a smooth upstream gradient and the reference's own bf16-autocast error as
    # calibration (round 1 only asked for a median cosine > 0.9 between L1 gradients, whose signs flip under bf16 noise)
    from util import gradient_parity_rows, assert_gradient_parity
    assert_gradient_parity(gradient_parity_rows(m, sd, cfg, x.cuda(), eps.cuda()), tag=str(flags))
