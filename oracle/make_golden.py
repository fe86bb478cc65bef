"""Generate tests/golden/*.pt from the UNMODIFIED reference (authoring container only).

For each fixture: weights = oracle.init_state_dict(cfg, seed, mode) (regenerable
anywhere from the seed), inputs = seeded torch.rand; outputs are produced by the
reference's own nn.Module tree imported from /root/reference.  Only the small
input/output tensors and a checksum of the weights are stored.

    python oracle/make_golden.py
"""
from __future__ import annotations

import hashlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import transvae_oracle as O  # noqa: E402
from validate_against_reference import (REF_MAIN, REF_PATCHED, build_reference_model,  # noqa: E402
                                        import_reference)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

FIXTURES = {
    # name: (cfg, seed, mode, batch, res)
    "mini_ref": (dict(depths=[1, 1, 1, 1, 2], base_dims=[64, 64, 64, 128, 128], mlp_ratio=1.0, head_dim=64,
                      latent_dim=32), 1, "reference", 2, 64),
    "mini_tamed": (dict(depths=[1, 1, 1, 1, 2], base_dims=[64, 64, 64, 128, 128], mlp_ratio=1.0, head_dim=64,
                        latent_dim=32), 2, "tamed", 2, 64),
    "mini_tamed_128": (dict(depths=[1, 1, 1, 1, 1], base_dims=[64, 64, 128, 128, 192], mlp_ratio=1.0, head_dim=64,
                            latent_dim=32), 3, "tamed", 1, 128),
}


def weights_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for name, (cfg, seed, mode, B, res) in FIXTURES.items():
        sd = O.init_state_dict(cfg, seed, mode)
        x = torch.rand(B, 3, res, res, generator=torch.Generator().manual_seed(1000 + seed))
        ref = build_reference_model(import_reference(REF_MAIN), cfg, sd)
        pkg_p = import_reference(REF_PATCHED)
        ref_p = build_reference_model(pkg_p, cfg, sd)
        with torch.no_grad():
            mu, logvar = ref.encode(x)
            recon_mu = ref.decode(mu)
            torch.manual_seed(77)
            recon_p, mu_p, logvar_p = ref_p(x)
            torch.manual_seed(77)
            eps = torch.randn(mu_p.shape)
        loss_fn = pkg_p.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)
        losses = loss_fn(recon_p, x, mu_p, logvar_p)
        # gradients of the patched training loss w.r.t. a few representative parameters
        ref_p.train()
        ref_p.zero_grad()
        torch.manual_seed(77)
        r2, m2, l2 = ref_p(x)
        loss_fn(r2, x, m2, l2)["total"].backward()
        gnames = ["encoder.conv_in.weight", "encoder.stages.0.0.norm1.weight", "encoder.stages.2.0.attn.to_q.weight",
                  "encoder.stages.2.0.attn.norm_k.bias", "encoder.stages.2.0.norm1.weight",
                  "encoder.stages.2.0.ffn.conv.2.weight", "encoder.downsamples.1.dc_conv.weight", "conv_mu.weight",
                  "decoder.conv_in.weight", "decoder.stages.0.0.ffn.proj_out.bias", "decoder.upsamples.0.dc_conv.weight",
                  "decoder.upsamples.3.main_path.1.weight", "decoder.stages.4.0.conv2.weight", "decoder.conv_out.weight"]
        # big tensors: keep the first 256 elements and the L2 norm only (fixtures stay small)
        named = dict(ref_p.named_parameters())
        grads = {k: dict(head=named[k].grad.flatten()[:256].clone(), norm=named[k].grad.norm().clone(),
                         numel=named[k].grad.numel()) for k in gnames}
        blob = dict(cfg=cfg, seed=seed, mode=mode, weights_sha256=weights_checksum(sd), x=x, mu=mu, logvar=logvar,
                    recon_from_mu=recon_mu, eps=eps, recon_patched=recon_p, mu_patched=mu_p, logvar_patched=logvar_p,
                    loss_l1=losses["l1"].detach(), loss_kl=losses["kl"].detach(), loss_total=losses["total"].detach(),
                    grads=grads, torch_version=torch.__version__)
        path = os.path.join(OUT, f"{name}.pt")
        torch.save(blob, path)
        sat_mu = float((mu_p.abs() >= 50).float().mean())
        sat_lv = float(((logvar_p <= -30) | (logvar_p >= 20)).float().mean())
        print(f"{name}: {os.path.getsize(path)/1024:.0f} KiB  |mu|max={float(mu.abs().max()):.3g} "
              f"|logvar|max={float(logvar.abs().max()):.3g} |recon|max={float(recon_mu.abs().max()):.3g} "
              f"sat(mu)={sat_mu:.2f} sat(logvar)={sat_lv:.2f} loss={float(losses['total']):.4f} "
              f"finite={bool(torch.isfinite(recon_p).all())}")


if __name__ == "__main__":
    main()
