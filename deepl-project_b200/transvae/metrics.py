"""On-GPU reconstruction metrics: the step after ``decode`` in the reference's evaluation scripts.

Mirrors ``calculate_psnr`` / ``calculate_ssim`` (evaluate_transvae.py:47-77), the per-image loop of ``evaluate_model``
(evaluate_transvae.py:110-176) and ``evaluate_resolution`` (scripts/reproduce/test_rope_extrapolation.py:28-51), but one
kernel launch (``tvae_metrics``) produces the per-image sums for a whole batch and nothing is copied to the host until
the caller asks for Python floats.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import ops

Tensor = torch.Tensor


def image_metrics(reconstruction: Tensor, images: Tensor, transform: Optional[str] = "sigmoid") -> Dict[str, Tensor]:
    """Per-image ``mse``, ``l1``, ``psnr`` (dB, max_val 1) and ``ssim`` as device tensors of shape [B].

    ``transform`` is applied to the reconstruction first: ``"sigmoid"`` (evaluate_transvae.py:131), ``"clamp"``
    (clamp(0, 1), evaluate.py:108-110) or ``None`` (raw, test_rope_extrapolation.py:44-47)."""
    acc = ops.metrics_sums(reconstruction, images, transform)
    n = reconstruction[0].numel()
    mse = acc[:, 0] / n
    psnr = torch.where(mse > 0, 20.0 * torch.log10(1.0 / torch.sqrt(mse.clamp_min(1e-45))), torch.full_like(mse, math.inf))
    return {"mse": mse, "l1": acc[:, 1] / n, "psnr": psnr, "ssim": acc[:, 2] / n}


def calculate_psnr(img1: Tensor, img2: Tensor, max_val: float = 1.0) -> float:
    """Reference signature (evaluate_transvae.py:47-53): PSNR over the whole tensor."""
    acc = ops.metrics_sums(img1, img2, None)
    mse = float(acc[:, 0].sum()) / img1.numel()
    return math.inf if mse == 0 else 20.0 * math.log10(max_val / math.sqrt(mse))


def calculate_ssim(img1: Tensor, img2: Tensor, window_size: int = 11, size_average: bool = True):
    """Reference signature (evaluate_transvae.py:56-77); only the reference's window of 11 is built."""
    if window_size != 11:
        raise NotImplementedError("the B200 metrics kernel implements the reference's 11x11 window only")
    acc = ops.metrics_sums(img1, img2, None)
    if size_average:
        return float(acc[:, 2].sum()) / img1.numel()
    return acc[:, 2] / img1[0].numel()


@torch.no_grad()
def evaluate_model(model, batches: Iterable, loss_fn=None, transform: Optional[str] = "sigmoid",
                   max_batches: Optional[int] = None) -> Tuple[Dict[str, float], Dict[str, float]]:
    """``evaluate_model`` of the reference (evaluate_transvae.py:110-176): averages / standard deviations of the
    per-image metrics over an iterable of image batches (tensors, or ``(images, label)`` pairs)."""
    model.eval()
    per = {"mse": [], "psnr": [], "ssim": []}
    comp = {"l1": [], "kl": []}
    for i, batch in enumerate(batches):
        if max_batches is not None and i >= max_batches:
            break
        images = batch[0] if isinstance(batch, (tuple, list)) else batch
        recon, mu, logvar = model(images)
        m = image_metrics(recon, images, transform)
        for k in per:
            per[k].append(m[k])
        if loss_fn is not None:
            losses = loss_fn(recon.float(), images.float(), mu.float(), logvar.float())
            comp["l1"].append(losses["l1"].detach().reshape(1))
            comp["kl"].append(losses["kl"].detach().reshape(1))
    avg, std = {}, {}
    for k, v in per.items():
        t = torch.cat(v).double()
        avg[k] = float(t.mean())
        std[k + "_std"] = float(t.std(unbiased=False))
    for k, v in comp.items():
        if v:
            avg[k] = float(torch.cat(v).double().mean())
    return avg, std
