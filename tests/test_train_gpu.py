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


@pytest.mark.parametrize("name", ["mini_tamed", "mini_ref"])
def test_whole_model_loss_and_gradients(name):
    blob, sd = load_golden(name)
    cfg = blob["cfg"]
    m = build_model(cfg, sd, patched=True).train()
    loss_fn = transvae.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)
    x = blob["x"].cuda()
    recon, mu, logvar = m(x, eps=blob["eps"].cuda())
    losses = loss_fn(recon, x, mu, logvar)
    losses["total"].backward()
    assert abs(float(losses["total"]) - float(blob["loss_total"])) < 5e-3
    # oracle gradients for every parameter (CPU fp32 autograd)
    sdg = {k: v.clone().requires_grad_("inv_freq" not in k) for k, v in sd.items()}
    rec_o, mu_o, lv_o, _ = O.forward(sdg, cfg, blob["x"], blob["eps"], patched=True)
    O.loss_l1_kl(rec_o, blob["x"], mu_o, lv_o, 1.0, 1e-8, patched=True)["total"].backward()
    worst_cos, worst_norm, n = 1.0, 0.0, 0
    report = {}
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        g, r = p.grad.detach().float().cpu().flatten(), sdg[k].grad.flatten()
        if float(r.norm()) < 1e-12:
            continue
        cos = float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-30))
        nr = abs(float(g.norm() / r.norm()) - 1.0)
        report[k] = (cos, nr)
        worst_cos, worst_norm, n = min(worst_cos, cos), max(worst_norm, nr), n + 1
    lo = sorted(report.items(), key=lambda kv: kv[1][0])[:5]
    print(name, "params:", n, "worst cosine:", worst_cos, "worst |norm ratio - 1|:", worst_norm, lo)
    if name == "mini_tamed":
        assert worst_cos > 0.99 and worst_norm < 0.05, lo
        # golden reference gradients (first 256 elements + norm)
        for k, gg in blob["grads"].items():
            g = dict(m.named_parameters())[k].grad.detach().float().cpu()
            assert abs(float(g.norm() / gg["norm"]) - 1) < 0.05, k
            assert rel(g.flatten()[:256], gg["head"]) < 0.1, (k, rel(g.flatten()[:256], gg["head"]))
    else:
        # reference init saturates the clamps: encoder gradients are ~0 in the reference too; the decoder's
        # must still agree
        dec = [v for k, v in report.items() if k.startswith("decoder.")]
        assert min(c for c, _ in dec) > 0.98


def test_gradient_checkpointing_matches():
    """T/test_installation.py:116-141: backward with gradient checkpointing enabled."""
    blob, sd = load_golden("mini_tamed")
    m = build_model(blob["cfg"], sd).train()
    x = blob["x"].cuda()
    eps = blob["eps"].cuda()
    loss_fn = transvae.TransVAELoss(lpips_weight=0.0, vf_weight=0.0, gan_weight=0.0)

    def grads():
        m.zero_grad()
        r, mu, lv = m(x, eps=eps)
        loss_fn(r, x, mu, lv)["total"].backward()
        return {k: p.grad.clone() for k, p in m.named_parameters()}

    g0 = grads()
    m.enable_gradient_checkpointing()
    g1 = grads()
    for k in g0:
        assert rel(g1[k], g0[k]) < 2e-2, k
