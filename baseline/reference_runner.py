"""Loads the UNMODIFIED reference (benabbouosama/DEEPL-Project, package `transvae`) from the git-ignored install under
baseline/_ref/ -- test / benchmark infrastructure, never imported by the product package.

Install recipe (run once in the authoring container; recorded in DESIGN.md):

    cp -r /root/reference/transvae-implementation /tmp/refsrc && cp /root/reference/README.md /tmp/refsrc/
    cp /root/reference/README.md /tmp/refsrc/transvae-implementation_patched/
    pip install --no-index --no-build-isolation --no-deps --target baseline/_ref/main    /tmp/refsrc
    pip install --no-index --no-build-isolation --no-deps --target baseline/_ref/patched /tmp/refsrc/transvae-implementation_patched

(setup.py reads a README.md that only exists one directory up, hence the copy; --no-deps because `lpips` / `timm` are
not in the wheelhouse.)  The two trees are loaded under the aliases `transvae_ref_main` / `transvae_ref_patched` so that
they can live next to this repository's own `transvae` package in one process; the reference only uses relative imports.
"""
from __future__ import annotations

import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

# transvae_large_f16d32.yaml:3-17 (`model:` block); the other variants follow TransVAE._get_variant_config
# (transvae/models/transvae.py:107-153)
VARIANTS = {
    "tiny": dict(depths=[2, 2, 2, 2, 2], base_dims=[64, 64, 128, 256, 512]),
    "base": dict(depths=[2, 2, 2, 3, 4], base_dims=[128, 128, 256, 512, 768]),
    "large": dict(depths=[3, 3, 3, 4, 6], base_dims=[192, 192, 384, 768, 1536]),
    "huge": dict(depths=[3, 3, 4, 6, 8], base_dims=[256, 256, 512, 1024, 2048]),
    "giant": dict(depths=[3, 3, 4, 8, 10], base_dims=[320, 320, 640, 1280, 2560]),
}


def available(patched: bool = True) -> bool:
    return os.path.exists(os.path.join(REF, "patched" if patched else "main", "transvae", "__init__.py"))


def load(patched: bool = True):
    """The reference package (module object) from baseline/_ref/{patched,main}, or None when it was not installed."""
    if not available(patched):
        return None
    alias = "transvae_ref_patched" if patched else "transvae_ref_main"
    if alias in sys.modules:
        return sys.modules[alias]
    stub = os.path.join(HERE, "lpips_stub")
    if "lpips" not in sys.modules and stub not in sys.path:
        sys.path.insert(0, stub)
    root = os.path.join(REF, "patched" if patched else "main", "transvae")
    spec = importlib.util.spec_from_file_location(alias, os.path.join(root, "__init__.py"), submodule_search_locations=[root])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


def model_config(variant: str = "large", latent_dim: int = 32) -> dict:
    v = VARIANTS[variant]
    return dict(variant=variant, compression_ratio=16, latent_dim=latent_dim, depths=list(v["depths"]),
                base_dims=list(v["base_dims"]), mlp_ratio=1.0, head_dim=64, use_rope=True, use_conv_ffn=True,
                use_dc_path=True)


def build(variant: str = "large", patched: bool = True, seed: int = 0, latent_dim: int = 32):
    """(reference TransVAE module, its package) with the reference's own random init under torch.manual_seed(seed)."""
    import torch
    pkg = load(patched)
    if pkg is None:
        return None, None
    cfg = model_config(variant, latent_dim)
    torch.manual_seed(seed)
    m = pkg.TransVAE(config=cfg, variant=variant, compression_ratio=16, latent_dim=latent_dim)
    return m, pkg
