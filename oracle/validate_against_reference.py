"""Pin the oracle against the UNMODIFIED reference (authoring container only).

Imports the reference packages from /root/reference (read-only) with a stub
``lpips`` module, loads the oracle's seeded weights into the reference's
``nn.Module`` tree via ``load_state_dict`` and checks, on CPU in fp32, that each
oracle function reproduces the reference bit-for-bit.  Run:

    python oracle/validate_against_reference.py

This script is test infrastructure; it cannot run on the GPU box
(/root/reference does not exist there) -- the committed fixtures under
tests/golden/ (written by oracle/make_golden.py) carry its result.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import transvae_oracle as O  # noqa: E402

REF_MAIN = "/root/reference/transvae-implementation"
REF_PATCHED = "/root/reference/transvae-implementation/transvae-implementation_patched"


def import_reference(root: str):
    """Import the reference's ``transvae`` package from ``root`` under a stub lpips."""
    if "lpips" not in sys.modules:
        stub = types.ModuleType("lpips")

        class LPIPS(torch.nn.Module):
            def __init__(self, net="vgg", **kw):
                super().__init__()

            def forward(self, a, b):
                return torch.zeros(a.shape[0], 1, 1, 1)

        stub.LPIPS = LPIPS
        sys.modules["lpips"] = stub
    for k in [k for k in sys.modules if k == "transvae" or k.startswith("transvae.")]:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        return importlib.import_module("transvae")
    finally:
        sys.path.remove(root)


def build_reference_model(pkg, cfg: dict, sd):
    model = pkg.TransVAE(config=cfg, variant="x", compression_ratio=16, latent_dim=cfg.get("latent_dim", 32))
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.eval()


MINI = dict(depths=[1, 1, 1, 1, 2], base_dims=[64, 64, 64, 128, 128], mlp_ratio=1.0, head_dim=64, latent_dim=32)


def check(name, a, b):
    same = torch.equal(a, b)
    err = float((a - b).abs().max())
    print(f"  {name:40s} equal={same} max|d|={err:.3e}")
    assert same or err == 0.0, name


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    for mode in ("reference", "tamed"):
        print(f"== mini config, init mode {mode}")
        sd = O.init_state_dict(MINI, seed=1, mode=mode)
        ref_main = build_reference_model(import_reference(REF_MAIN), MINI, sd)
        # key set / shapes / order must match exactly
        rsd = ref_main.state_dict()
        assert list(rsd.keys()) == list(sd.keys()), "state_dict key order differs"
        for k in sd:
            assert tuple(rsd[k].shape) == tuple(sd[k].shape), k
        check("inv_freq", rsd["encoder.stages.2.0.attn.rope.inv_freq"], sd["encoder.stages.2.0.attn.rope.inv_freq"])
        x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(5))
        with torch.no_grad():
            mu_r, lv_r = ref_main.encode(x)
            rec_r = ref_main.decode(mu_r)
            mu_o, lv_o = O.encode(sd, MINI, x)
            rec_o = O.decode(sd, MINI, mu_o)
        check("encode.mu", mu_o, mu_r)
        check("encode.logvar", lv_o, lv_r)
        check("decode(mu)", rec_o, rec_r)

        # per-module checks
        blk = ref_main.encoder.stages[2][0]
        h = torch.randn(2, 64, 16, 16, generator=torch.Generator().manual_seed(6)) * 3
        with torch.no_grad():
            check("TransVAEBlock", O.transvae_block(sd, "encoder.stages.2.0.", h, 64), blk(h))
            check("RMSNorm", O.rmsnorm(h, sd["encoder.stages.2.0.norm1.weight"]), blk.norm1(h))
            check("attention", O.attention(sd, "encoder.stages.2.0.attn.", h, 64), blk.attn(h))
            check("ConvFFN", O.conv_ffn(sd, "encoder.stages.2.0.ffn.", h), blk.ffn(h))
            check("ResBlock", O.resblock(sd, "encoder.stages.0.0.", h), ref_main.encoder.stages[0][0](h))
            check("Downsample", O.downsample(sd, "encoder.downsamples.2.", h), ref_main.encoder.downsamples[2](h))
            h2 = torch.randn(2, 128, 8, 8, generator=torch.Generator().manual_seed(7))
            check("Upsample", O.upsample(sd, "decoder.upsamples.1.", h2), ref_main.decoder.upsamples[1](h2))
            q = torch.randn(2, 2, 48, 64, generator=torch.Generator().manual_seed(8))
            check("RoPE2D", O.rope2d(q, 6, 8, sd["encoder.stages.2.0.attn.rope.inv_freq"]), blk.attn.rope(q, 6, 8))
            ang = O.rope_angles(6, 8, sd["encoder.stages.2.0.attn.rope.inv_freq"])
            cs, sn = ang.cos(), ang.sin()
            qe, qo = q[..., 0::2], q[..., 1::2]
            closed = torch.stack([qe * cs[:, 0::2] - qo * sn[:, 0::2], qe * sn[:, 1::2] + qo * cs[:, 1::2]], -1).flatten(-2)
            check("RoPE closed form", closed, blk.attn.rope(q, 6, 8))

        # patched forward (clamps + fp32 reparam) and both loss flavours
        ref_p_pkg = import_reference(REF_PATCHED)
        ref_p = build_reference_model(ref_p_pkg, MINI, sd)
        with torch.no_grad():
            torch.manual_seed(11)
            rec_r, mu_r, lv_r = ref_p(x)
            torch.manual_seed(11)
            eps = torch.randn(mu_r.shape)
            rec_o, mu_o, lv_o, _ = O.forward(sd, MINI, x, eps, patched=True)
        check("patched forward recon", rec_o, rec_r)
        check("patched forward mu", mu_o, mu_r)
        check("patched forward logvar", lv_o, lv_r)
        loss_p = ref_p_pkg.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)
        lr = loss_p(rec_r, x, mu_r, lv_r)
        lo = O.loss_l1_kl(rec_o, x, mu_o, lv_o, 1.0, 1e-8, patched=True)
        check("patched loss l1", lo["l1"], lr["l1"])
        check("patched loss kl", lo["kl"], lr["kl"])
        check("patched loss total", lo["total"], lr["total"])
        main_pkg = import_reference(REF_MAIN)
        loss_m = main_pkg.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)
        lr = loss_m(rec_r, x, mu_r, lv_r)
        lo = O.loss_l1_kl(rec_o, x, mu_o, lv_o, 1.0, 1e-8, patched=False)
        check("main loss l1", lo["l1"], lr["l1"])
        check("main loss kl", lo["kl"], lr["kl"])

        # gradients through the patched path
        ref_p.train()
        xg = x.clone()
        torch.manual_seed(12)
        rec_r, mu_r, lv_r = ref_p(xg)
        loss_p(rec_r, xg, mu_r, lv_r)["total"].backward()
        sdg = {k: v.clone().requires_grad_(v.is_floating_point() and "inv_freq" not in k) for k, v in sd.items()}
        torch.manual_seed(12)
        eps = torch.randn(mu_r.shape)
        rec_o, mu_o, lv_o, _ = O.forward(sdg, MINI, xg, eps, patched=True)
        O.loss_l1_kl(rec_o, xg, mu_o, lv_o, 1.0, 1e-8, patched=True)["total"].backward()
        worst = 0.0
        for k, p in ref_p.named_parameters():
            g = sdg[k].grad
            assert g is not None, k
            worst = max(worst, float((g - p.grad).abs().max()))
        print(f"  gradient parity over {len(list(ref_p.named_parameters()))} tensors: max|d|={worst:.3e}")
        assert worst == 0.0

    # ablation switches of the constructor (transvae.py:36-38): use_rope=False and use_dc_path=False change the module
    # tree and must stay bit-identical; use_conv_ffn=False is broken in the reference itself (nn.Linear along W of NCHW)
    for flags in (dict(use_rope=False), dict(use_dc_path=False), dict(use_rope=False, use_dc_path=False)):
        cfg = dict(MINI, **flags)
        sda = O.init_state_dict(cfg, seed=4, mode="tamed")
        ref = import_reference(REF_MAIN).TransVAE(config=MINI, variant="x", compression_ratio=16, latent_dim=32, **flags).eval()
        assert list(ref.state_dict().keys()) == list(sda.keys()), flags
        ref.load_state_dict(sda, strict=True)
        xa = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(6))
        with torch.no_grad():
            mu_r, lv_r = ref.encode(xa)
            rec_r = ref.decode(mu_r)
            mu_o, lv_o = O.encode(sda, cfg, xa)
            rec_o = O.decode(sda, cfg, mu_o)
        check(f"ablation {flags} mu", mu_o, mu_r)
        check(f"ablation {flags} logvar", lv_o, lv_r)
        check(f"ablation {flags} recon", rec_o, rec_r)
    try:
        bad = import_reference(REF_MAIN).TransVAE(config=MINI, variant="x", compression_ratio=16, latent_dim=32, use_conv_ffn=False)
        with torch.no_grad():
            bad.encode(torch.rand(1, 3, 64, 64))
        raise AssertionError("the reference was expected to fail with use_conv_ffn=False")
    except RuntimeError as e:
        print(f"  use_conv_ffn=False: the reference raises RuntimeError ({str(e)[:60]}...) -- nothing to mirror")

    # module-level variants reachable by constructing the reference's modules directly: ConvFFN(conv_type='depthwise')
    # (conv.py:42-50) and ResBlock with a convolutional shortcut (blocks.py:40-46), forward and parameter gradients
    ref_pkg = import_reference(REF_MAIN)
    conv_mod = importlib.import_module("transvae.modules.conv")
    blocks_mod = importlib.import_module("transvae.modules.blocks")
    torch.manual_seed(21)
    cases = [("ConvFFN depthwise", conv_mod.ConvFFN(128, conv_type="depthwise"), lambda s, t: O.conv_ffn(s, "", t), (2, 128, 8, 8)),
             ("ResBlock 1x1 shortcut", blocks_mod.ResBlock(64, 128), lambda s, t: O.resblock(s, "", t), (2, 64, 8, 8)),
             ("ResBlock 3x3 shortcut", blocks_mod.ResBlock(64, 128, use_conv_shortcut=True), lambda s, t: O.resblock(s, "", t),
              (2, 64, 8, 8))]
    for name, mod, fn, shape in cases:
        sdm = {k: v.detach().clone() for k, v in mod.state_dict().items()}
        xm = torch.randn(*shape, generator=torch.Generator().manual_seed(22))
        out_r = mod(xm)
        gout = torch.randn(out_r.shape, generator=torch.Generator().manual_seed(23))
        out_r.backward(gout)
        sdg = {k: v.clone().requires_grad_(True) for k, v in sdm.items()}
        out_o = fn(sdg, xm)
        out_o.backward(gout)
        check(name, out_o.detach(), out_r.detach())
        worst = max(float((sdg[k].grad - p.grad).abs().max()) for k, p in mod.named_parameters())
        print(f"  {name}: gradient parity over {len(sdm)} tensors max|d|={worst:.3e}")
        assert worst == 0.0
    del ref_pkg

    # parameter counts and FLOP model vs SURVEY section 6 [measured] numbers
    for v, (f, d), want in [("tiny", (16, 32), 81.9e6), ("large", (16, 32), 1049.2e6), ("giant", (16, 32), 4837.3e6)]:
        n = O.count_params(O.variant_config(v, f, d))["total"]
        print(f"  params {v}: {n/1e6:.1f} M (survey {want/1e6:.1f} M)")
        assert abs(n - want) / want < 1e-3
    for res, want in [(256, 2062.6), (512, 10473.0), (1024, 77455.0)]:
        g = O.forward_flops_per_image(O.variant_config("large"), res) / 1e9
        print(f"  large fwd GFLOP/img @{res}: {g:.1f} (survey {want})")
        assert abs(g - want) / want < 2e-3
    print("oracle == reference: OK")


if __name__ == "__main__":
    main()
