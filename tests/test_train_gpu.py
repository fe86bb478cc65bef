"""Gradient parity of the training path (forward + hand-written backward kernels) on a real B200.

Per block: dx and every parameter gradient against torch.autograd through the oracle (fp32 CPU) for the same block,
same seeded weights / inputs / upstream gradient: max|ours-ref|/max|ref| <= 3e-2 (bf16 activations and activation
gradients, fp32 accumulation).  Whole model: the patched training loss (L1 + 1e-8 KL) and its gradients against the
golden reference values and the oracle: per-tensor cosine >= 0.99, norm within 5 %.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

import transvae  # noqa: E402
import transvae_oracle as O  # noqa: E402
from util import build_model, load_golden, nchw_f32, nhwc_bf16, rel  # noqa: E402

TOL = 3e-2


def oracle_block_grads(fn, sd, prefix, x, dout):
    keys = [k for k in sd if k.startswith(prefix) and "inv_freq" not in k]
    sdg = dict(sd)
    for k in keys:
        sdg[k] = sd[k].clone().requires_grad_(True)
    xg = x.clone().requires_grad_(True)
    out = fn(sdg, xg)
    out.backward(dout)
    return out.detach(), xg.grad, {k: sdg[k].grad for k in keys}


def run_block(mod, x_nchw, dout_nchw):
    mod.train()
    mod.zero_grad()
    x = nhwc_bf16(x_nchw).cuda().requires_grad_(True)
    out = mod.forward_nhwc(x)
    out.backward(nhwc_bf16(dout_nchw).cuda())
    return nchw_f32(out.detach()), nchw_f32(x.grad)


@pytest.mark.parametrize("kind", ["resblock", "transvae_block", "downsample", "upsample"])
def test_block_gradients(kind):
    blob, sd = load_golden("mini_tamed")
    cfg = blob["cfg"]
    m = build_model(cfg, sd)
    g = torch.Generator().manual_seed(11)
    if kind == "resblock":
        prefix, mod = "encoder.stages.0.0.", m.encoder.stages[0][0]
        x = torch.randn(2, 64, 32, 32, generator=g) * 2
        fn = lambda s, t: O.resblock(s, prefix, t)
    elif kind == "transvae_block":
        prefix, mod = "encoder.stages.3.0.", m.encoder.stages[3][0]
        x = torch.randn(2, 128, 8, 8, generator=g) * 2
        fn = lambda s, t: O.transvae_block(s, prefix, t, 64)
    elif kind == "downsample":
        prefix, mod = "encoder.downsamples.2.", m.encoder.downsamples[2]
        x = torch.randn(2, 64, 16, 16, generator=g) * 2
        fn = lambda s, t: O.downsample(s, prefix, t)
    else:
        prefix, mod = "decoder.upsamples.1.", m.decoder.upsamples[1]
        x = torch.randn(2, 128, 8, 8, generator=g) * 2
        fn = lambda s, t: O.upsample(s, prefix, t)
    x = x.to(torch.bfloat16).float()
    out_ref, _, _ = oracle_block_grads(fn, sd, prefix, x, torch.zeros(1).expand_as(fn(sd, x)).clone())
    dout = (torch.randn(out_ref.shape, generator=g)).to(torch.bfloat16).float()
    out_ref, dx_ref, gref = oracle_block_grads(fn, sd, prefix, x, dout)
    out, dx = run_block(mod, x, dout)
    errs = {"out": rel(out, out_ref), "dx": rel(dx, dx_ref)}
    for name, p in mod.named_parameters():
        errs[name] = rel(p.grad, gref[prefix + name])
    print(kind, {k: round(v, 4) for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if v > TOL}
    assert not bad, bad


from util import smooth_objective as _smooth_objective  # noqa: E402
from util import gradient_parity_rows, assert_gradient_parity  # noqa: E402


@pytest.mark.parametrize("name", ["mini_tamed", "mini_tamed_128"])
def test_whole_model_gradients_vs_oracle_with_bf16_calibration(name):
    blob, sd = load_golden(name)
    m = build_model(blob["cfg"], sd, patched=True).train()
    rows = gradient_parity_rows(m, sd, blob["cfg"], blob["x"].cuda(), blob["eps"].cuda())
    assert_gradient_parity(rows, tag=name)


@pytest.mark.parametrize("name", ["mini_tamed", "mini_ref"])
def test_training_loss_matches_golden(name):
    blob, sd = load_golden(name)
    m = build_model(blob["cfg"], sd, patched=True).train()
    loss_fn = transvae.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)
    x = blob["x"].cuda()
    recon, mu, logvar = m(x, eps=blob["eps"].cuda())
    losses = loss_fn(recon, x, mu, logvar)
    losses["total"].backward()
    # bf16 forward against the reference's fp32 loss.  The tamed fixture agrees to 1e-4; the untamed random init
    # ("mini_ref": activations grow to 1e2, every block amplifies the bf16 rounding of the one before) sits at
    # 0.30 - 0.33 % of its loss of 1.585, so its bar is 1e-2 absolute (0.6 %) rather than a coin flip at 5e-3
    tol = 5e-3 if name == "mini_tamed" else 1e-2
    assert abs(float(losses["total"].detach()) - float(blob["loss_total"])) < tol
    assert abs(float(losses["l1"].detach()) - float(blob["loss_l1"])) < tol
    for k, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k


def test_gradient_checkpointing_matches():
    """T/test_installation.py:116-141: backward with gradient checkpointing enabled.  Forward and backward are
    bit-reproducible (fixed-order reductions everywhere: test_backward_is_bit_reproducible), and the recomputed forward
    runs the same kernels on the same inputs -- so the checkpointed backward must reproduce the plain one EXACTLY, per
    block and for the whole model.  (Round 1 had atomics in the weight-gradient flush, the GroupNorm / token-norm / bias
    reductions and the attention dQ accumulation: two identical plain runs differed by 9 % median l2 per tensor.)"""
    import torch.utils.checkpoint as cp
    blob, sd = load_golden("mini_tamed")
    m = build_model(blob["cfg"], sd).train()
    l2 = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))
    for blk, shape in ((m.encoder.stages[0][0], (2, 32, 32, 64)), (m.encoder.stages[3][0], (2, 8, 8, 128)),
                       (m.encoder.downsamples[2], (2, 16, 16, 64)), (m.decoder.upsamples[1], (2, 8, 8, 128))):
        xb = (torch.randn(shape, generator=torch.Generator().manual_seed(3)) * 2).to(torch.bfloat16).cuda()
        res, dout = [], None
        for use in (False, True):
            blk.zero_grad()
            xi = xb.clone().requires_grad_(True)
            out = cp.checkpoint(blk.forward_nhwc, xi, use_reentrant=False) if use else blk.forward_nhwc(xi)
            if dout is None:
                dout = torch.randn(out.shape, generator=torch.Generator().manual_seed(4)).to(torch.bfloat16).cuda()
            out.backward(dout)
            res.append((xi.grad.clone(), {k: p.grad.clone() for k, p in blk.named_parameters()}))
        assert torch.equal(res[1][0], res[0][0]), l2(res[1][0], res[0][0])
        for k in res[0][1]:
            assert torch.equal(res[1][1][k], res[0][1][k]), (k, l2(res[1][1][k], res[0][1][k]))
    # whole model through model.enable_gradient_checkpointing()
    x, eps = blob["x"].cuda(), blob["eps"].cuda()
    g = torch.Generator().manual_seed(5)
    G = [torch.randn(blob["x"].shape, generator=g).cuda(), torch.randn(blob["mu"].shape, generator=g).cuda(),
         torch.randn(blob["mu"].shape, generator=g).cuda()]

    def grads():
        m.zero_grad()
        r, mu, lv = m(x, eps=eps)
        _smooth_objective(r, mu, lv, G).backward()
        return {k: p.grad.clone() for k, p in m.named_parameters()}

    g0 = grads()
    m.enable_gradient_checkpointing()
    g1 = grads()
    diff = {k: l2(g1[k], g0[k]) for k in g0 if not torch.equal(g1[k], g0[k])}
    assert not diff, sorted(diff.items(), key=lambda kv: -kv[1])[:5]


@pytest.mark.parametrize("name", ["mini_tamed", "mini_tamed_128"])
def test_backward_is_bit_reproducible(name):
    """Two forward + backward passes on the same inputs give bit-identical losses and parameter gradients: no reduction of
    the training step depends on the order in which thread blocks finish (weight-gradient pixel splits, GroupNorm /
    token-norm / bias column sums, loss sums, attention dQ accumulation)."""
    blob, sd = load_golden(name)
    m = build_model(blob["cfg"], sd, patched=True).train()
    loss_fn = transvae.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)
    x, eps = blob["x"].cuda(), blob["eps"].cuda()
    runs = []
    for _ in range(3):
        m.zero_grad()
        recon, mu, logvar = m(x, eps=eps)
        losses = loss_fn(recon, x, mu, logvar)
        losses["total"].backward()
        runs.append((losses["total"].detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()}))
    for r in runs[1:]:
        assert torch.equal(r[0], runs[0][0])
        bad = [k for k in runs[0][1] if not torch.equal(r[1][k], runs[0][1][k])]
        assert not bad, bad[:8]
