"""Shared helpers for the GPU parity tests (test infrastructure)."""
import hashlib
import os

import torch

import transvae
import transvae_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def weights_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def load_golden(name):
    blob = torch.load(os.path.join(GOLDEN, f"{name}.pt"), map_location="cpu", weights_only=False)
    sd = O.init_state_dict(blob["cfg"], blob["seed"], blob["mode"])
    assert weights_checksum(sd) == blob["weights_sha256"], "seeded weights differ from the ones the golden was made with"
    return blob, sd


def build_model(cfg, sd, device="cuda", patched=True):
    with torch.device("meta"):
        m = transvae.TransVAE(config=cfg, latent_dim=cfg.get("latent_dim", 32), patched=patched,
                              use_rope=cfg.get("use_rope", True), use_conv_ffn=cfg.get("use_conv_ffn", True),
                              use_dc_path=cfg.get("use_dc_path", True))
    m = m.to_empty(device=device)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def nhwc_bf16(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw_f32(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def rel(ours, ref):
    return O.max_rel_err(ours.detach().cpu(), ref.detach().cpu())
