"""TransVAE -- B200-native drop-in for transvae/models/transvae.py.

Same constructor (both the documented README form and the ``config=`` dict form the reference's scripts use --
SURVEY fact 1), same ``state_dict`` keys, same methods: ``encode``, ``reparameterize``, ``decode``, ``forward``
-> ``(reconstruction, mu, logvar)``, ``get_last_layer``, ``from_pretrained``, ``enable_gradient_checkpointing``,
``get_num_params``.  ``patched=True`` selects the numerically safe semantics of the reference's
transvae-implementation_patched tree (clamps at transvae.py:186-196, :244-245).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import _taps as T
from .. import kernels as K
from .._pack import PackCache, bf16c, f32c
from ..modules.blocks import _Conv2dParams
from .decoder import TransVAEDecoder
from .encoder import TransVAEEncoder

_VARIANTS = {
    "tiny_f16d32": {"depths": [3, 3, 3, 3, 3], "base_dims": [128, 128, 256, 256, 512]},
    "base_f16d32": {"depths": [3, 3, 3, 3, 3], "base_dims": [128, 128, 256, 512, 1024]},
    "large_f16d32": {"depths": [3, 3, 3, 4, 6], "base_dims": [192, 192, 384, 768, 1536]},
    "huge_f16d32": {"depths": [3, 3, 4, 6, 8], "base_dims": [256, 256, 512, 1024, 2048]},
    "giant_f16d32": {"depths": [3, 3, 4, 8, 10], "base_dims": [320, 320, 640, 1280, 2560]},
    "large_f8d16": {"depths": [3, 3, 6, 8], "base_dims": [192, 384, 768, 1536]},
}


class TransVAE(nn.Module):
    def __init__(self, config: Optional[dict] = None, variant: str = "large", compression_ratio: int = 16,
                 latent_dim: int = 32, input_channels: int = 3, use_rope: bool = True, use_conv_ffn: bool = True,
                 use_dc_path: bool = True, patched: bool = True, **kwargs):
        super().__init__()
        # kwargs swallows input_resolution (README form), like the reference's **kwargs (transvae.py:37)
        self.variant, self.compression_ratio, self.latent_dim = variant, compression_ratio, latent_dim
        self.patched = patched
        if config is None:
            config = self._get_variant_config(variant, compression_ratio, latent_dim)
        depths, dims = list(config.get("depths")), list(config.get("base_dims"))
        mlp_ratio, head_dim = config.get("mlp_ratio", 1.0), config.get("head_dim", 64)
        self.encoder = TransVAEEncoder(input_channels=input_channels, latent_dim=latent_dim, depths=depths,
                                       base_dims=dims, compression_ratio=compression_ratio, mlp_ratio=mlp_ratio,
                                       head_dim=head_dim, use_rope=use_rope, use_conv_ffn=use_conv_ffn,
                                       use_dc_path=use_dc_path)
        self.conv_mu = _Conv2dParams(dims[-1], latent_dim, 3)
        self.conv_logvar = _Conv2dParams(dims[-1], latent_dim, 3)
        self.decoder = TransVAEDecoder(latent_dim=latent_dim, output_channels=input_channels, depths=depths[::-1],
                                       base_dims=dims[::-1], compression_ratio=compression_ratio, mlp_ratio=mlp_ratio,
                                       head_dim=head_dim, use_rope=use_rope, use_conv_ffn=use_conv_ffn,
                                       use_dc_path=use_dc_path)
        object.__setattr__(self, "_packs", PackCache())
        self._initialize_weights()

    def _apply(self, fn, *a, **kw):
        self._packs.clear()
        return super()._apply(fn, *a, **kw)

    @staticmethod
    def _get_variant_config(variant: str, f: int, d: int) -> dict:
        key = f"{variant}_f{f}d{d}"
        if key not in _VARIANTS:
            raise ValueError(f"Unknown variant: {variant} with f{f}d{d}")
        cfg = dict(_VARIANTS[key])
        cfg.update(mlp_ratio=1.0, head_dim=64)
        return cfg

    def _initialize_weights(self):
        """transvae.py:155-168: Kaiming-normal(fan_out, relu) convs, trunc-normal(0.02) linears, zero biases,
        unit norm scales (RMSNorm weights stay ones)."""
        from ..modules.blocks import _LinearParams, _NormParams
        for m in self.modules():
            if isinstance(m, _Conv2dParams):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, _LinearParams):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, _NormParams):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0)

    # ------------------------------------------------------------------------------------------
    def _heads(self, h: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """conv_mu and conv_logvar (transvae.py:182-183) as ONE 3x3 implicit GEMM with N = 2*latent_dim."""
        B, H, W, C = h.shape
        d = self.latent_dim
        npad = (2 * d + 63) // 64 * 64
        if K.needs_grad(h, self.conv_mu.weight, self.conv_logvar.weight):
            from .._autograd import HeadFn
            both = HeadFn.apply(h, T.pack_conv3x3(torch.cat([self.conv_mu.weight, self.conv_logvar.weight], 0), cout_pad=npad),
                                torch.nn.functional.pad(torch.cat([self.conv_mu.bias, self.conv_logvar.bias]), (0, npad - 2 * d)),
                                2 * d)
            return both[:, :d], both[:, d:]
        w = self._packs.get("heads", [self.conv_mu.weight, self.conv_logvar.weight], lambda: bf16c(
            T.pack_conv3x3(torch.cat([self.conv_mu.weight, self.conv_logvar.weight], 0), cout_pad=npad)))
        b = self._packs.get("heads_b", [self.conv_mu.bias, self.conv_logvar.bias], lambda: f32c(
            torch.nn.functional.pad(torch.cat([self.conv_mu.bias, self.conv_logvar.bias]), (0, npad - 2 * d))))
        both = K.mtgemm(T.plan_conv3x3(C), h, w, bias=b, out_f32_shape=(B, 2 * d, H, W), out_n=2 * d)
        return both[:, :d].contiguous(), both[:, d:].contiguous()

    def encode(self, x: torch.Tensor, trace: dict = None) -> Tuple[torch.Tensor, torch.Tensor]:
        h = self.encoder.forward_features(x, trace)
        return self._heads(h)

    def reparameterize(self, mu: torch.Tensor, logvar: torch.Tensor, eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        if eps is None:
            eps = torch.randn_like(mu, dtype=torch.float32)
        lv = logvar.clamp(-30.0, 20.0) if self.patched else logvar      # patched :189 clamps logvar only
        z, _, _ = K.reparam(mu, lv, eps, False)
        return z.to(mu.dtype)

    def decode(self, z: torch.Tensor, trace: dict = None) -> torch.Tensor:
        return self.decoder(z, trace) if trace is not None else self.decoder(z)

    def forward(self, x: torch.Tensor, return_dict: bool = False, eps: Optional[torch.Tensor] = None):
        mu, logvar = self.encode(x)
        if eps is None:
            eps = torch.randn_like(mu, dtype=torch.float32)
        z, mu, logvar = K.reparam(mu, logvar, eps, self.patched)
        reconstruction = self.decode(z)
        if return_dict:
            return {"reconstruction": reconstruction, "mu": mu, "logvar": logvar, "z": z}
        return reconstruction, mu, logvar

    def get_last_layer(self):
        return self.decoder.conv_out.weight

    @classmethod
    def from_pretrained(cls, model_name: str, **kwargs):
        variant, config = model_name.split("-")[1:3]
        f, d = int(config[1:].split("d")[0]), int(config.split("d")[1])
        return cls(variant=variant, compression_ratio=f, latent_dim=d, **kwargs)

    def enable_gradient_checkpointing(self):
        self.encoder.enable_gradient_checkpointing()
        self.decoder.enable_gradient_checkpointing()

    def get_num_params(self) -> dict:
        enc = sum(p.numel() for p in self.encoder.parameters())
        dec = sum(p.numel() for p in self.decoder.parameters())
        return {"encoder": enc, "decoder": dec, "total": sum(p.numel() for p in self.parameters())}


def create_transvae(variant: str = "large", compression_ratio: int = 16, latent_dim: int = 32, **kwargs) -> TransVAE:
    return TransVAE(variant=variant, compression_ratio=compression_ratio, latent_dim=latent_dim, **kwargs)
