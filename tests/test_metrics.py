"""Evaluation metrics: the oracle restatement against the reference's own outputs (CPU) and the
tvae_metrics kernel against both (GPU).  tests/golden/metrics_ref.pt is produced by oracle/make_golden_metrics.py from
the unmodified reference functions (evaluate_transvae.py:47-77)."""
import math
import os

import pytest
import torch

import transvae_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_ref.pt")
XFORM = {"sigmoid": torch.sigmoid, "clamp": lambda t: t.clamp(0, 1), "none": lambda t: t}


def _cases():
    return torch.load(GOLDEN, map_location="cpu", weights_only=False)["cases"]


def test_oracle_matches_reference_metrics():
    for name, case in _cases().items():
        for mode, f in XFORM.items():
            r, t, want = f(case["logits"]), case["target"], case["out"][mode]
            for i in range(r.shape[0]):
                assert abs(O.psnr(r[i:i + 1], t[i:i + 1]) - float(want["psnr"][i])) < 1e-5, (name, mode, i)
                assert abs(O.ssim(r[i:i + 1], t[i:i + 1]) - float(want["ssim"][i])) < 1e-6, (name, mode, i)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["sigmoid", "clamp", "none"])
def test_metrics_kernel_matches_reference(mode):
    from transvae import metrics
    for name, case in _cases().items():
        logits, target, want = case["logits"].cuda(), case["target"].cuda(), case["out"][mode]
        got = metrics.image_metrics(logits, target, None if mode == "none" else mode)
        # fp32 sums in a different order: 1e-4 relative on the means, 2e-3 dB on PSNR (tolerances of this test)
        assert torch.allclose(got["mse"].cpu(), want["mse"], rtol=1e-4, atol=1e-7), (name, mode)
        assert torch.allclose(got["l1"].cpu(), want["l1"], rtol=1e-4, atol=1e-7), (name, mode)
        assert float((got["psnr"].cpu() - want["psnr"]).abs().max()) < 2e-3, (name, mode)
        assert float((got["ssim"].cpu() - want["ssim"]).abs().max()) < 2e-4, (name, mode)


@pytest.mark.gpu
def test_metrics_reference_signatures_and_full_size():
    from transvae import metrics
    g = torch.Generator().manual_seed(3)
    t = torch.rand(4, 3, 256, 256, generator=g)
    r = (t + 0.05 * torch.randn(4, 3, 256, 256, generator=g)).clamp(0, 1)
    assert abs(metrics.calculate_psnr(r.cuda(), t.cuda()) - O.psnr(r, t)) < 2e-3
    assert abs(metrics.calculate_ssim(r.cuda(), t.cuda()) - O.ssim(r, t)) < 2e-4
    per = metrics.calculate_ssim(r.cuda(), t.cuda(), size_average=False).cpu()
    assert torch.allclose(per, O.ssim(r, t, size_average=False), atol=2e-4)
    assert metrics.calculate_psnr(t.cuda(), t.cuda()) == math.inf            # identical images (evaluate_transvae.py:50-51)
    with pytest.raises(NotImplementedError):
        metrics.calculate_ssim(r.cuda(), t.cuda(), window_size=7)
