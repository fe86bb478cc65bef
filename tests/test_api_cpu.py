"""Boundary checks that need no GPU: the reference's six installation checks (T/test_installation.py:10-175,
restated), state_dict compatibility with the reference, C-ABI symbol export, loud failure without a device."""
import ctypes
import os
import re

import pytest
import torch

import transvae
import transvae_oracle as O
from transvae import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINI = dict(depths=[1, 1, 1, 1, 2], base_dims=[64, 64, 64, 128, 128], mlp_ratio=1.0, head_dim=64)


def test_exports():
    assert set(transvae.__all__) == {"TransVAE", "create_transvae", "TransVAELoss"}
    assert callable(transvae.create_transvae)


@pytest.mark.parametrize("variant", ["tiny", "base"])
def test_readme_constructor_and_param_counts(variant):
    # README.md:102-107 form (no config dict), input_resolution swallowed
    m = transvae.TransVAE(variant=variant, compression_ratio=16, latent_dim=32, input_resolution=256)
    n = m.get_num_params()
    want = O.count_params(O.variant_config(variant))
    assert n == want
    assert m.variant == variant and m.compression_ratio == 16 and m.latent_dim == 32


def test_large_param_count_on_meta():
    with torch.device("meta"):
        m = transvae.TransVAE(variant="large", compression_ratio=16, latent_dim=32)
    assert abs(m.get_num_params()["total"] - 1049.2e6) < 0.1e6


def test_config_constructor_and_state_dict_keys_match_reference():
    m = transvae.TransVAE(config=MINI, latent_dim=32)
    ours = m.state_dict()
    ref = O.param_shapes(MINI)
    assert list(ours.keys()) == [k for k, _, _ in ref]
    for k, shape, _ in ref:
        assert tuple(ours[k].shape) == tuple(shape), k
    sd = O.init_state_dict(MINI, seed=3)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.state_dict()["decoder.stages.0.1.ffn.conv.2.weight"], sd["decoder.stages.0.1.ffn.conv.2.weight"])
    assert m.get_last_layer() is m.decoder.conv_out.weight


def test_f8_variant_has_four_stages():
    with torch.device("meta"):
        m = transvae.TransVAE(variant="large", compression_ratio=8, latent_dim=16)
    assert len(m.encoder.stages) == 4 and len(m.encoder.downsamples) == 3


def test_unknown_variant_raises():
    with pytest.raises(ValueError):
        transvae.TransVAE(variant="nope", compression_ratio=16, latent_dim=32)


def test_from_pretrained_parses_name():
    with torch.device("meta"):
        m = transvae.TransVAE.from_pretrained("transvae-tiny-f16d32")
    assert m.variant == "tiny" and m.compression_ratio == 16 and m.latent_dim == 32


def test_init_statistics_follow_reference_init():
    torch.manual_seed(0)
    m = transvae.TransVAE(config=MINI, latent_dim=32)
    sd = m.state_dict()
    w = sd["encoder.stages.0.0.conv1.weight"]
    assert abs(float(w.std()) - (2.0 / (64 * 9)) ** 0.5) < 5e-3
    assert float(sd["encoder.stages.0.0.conv1.bias"].abs().max()) == 0
    lw = sd["encoder.stages.2.0.attn.to_q.weight"]
    assert abs(float(lw.std()) - 0.02) < 2e-3 and float(lw.abs().max()) <= 2.0
    assert torch.equal(sd["encoder.stages.2.0.norm1.weight"], torch.ones(64))


def test_loss_rejects_out_of_scope_terms():
    with pytest.raises(NotImplementedError):
        transvae.TransVAELoss()                       # reference default lpips_weight=1.0 needs VGG weights
    transvae.TransVAELoss(lpips_weight=0.0, vf_weight=0.0, gan_weight=0.0)


def test_cpu_tensors_fail_loudly():
    m = transvae.TransVAE(config=MINI, latent_dim=32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.encode(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            m.decode(torch.zeros(1, 32, 4, 4))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert lib.tvae_abi_version() == 4
    hdr = open(os.path.join(ROOT, "include", "transvae_sm100.h")).read()
    declared = set(re.findall(r"\b(tvae_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/transvae_sm100.h but not exported"
    assert declared == set(_lib.exported_symbols())


def test_compute_entry_points_refuse_without_device():
    if torch.cuda.is_available():
        pytest.skip("device present")
    lib = _lib.load()
    assert lib.tvae_device_ok() == 0
    rc = lib.tvae_attn_fwd(None, None, None, 1, 16, 64, None)
    assert rc != 0 and "no CPU fallback" in _lib.last_error()


@pytest.mark.parametrize("flags", [dict(use_rope=False), dict(use_dc_path=False), dict(use_rope=False, use_dc_path=False)])
def test_ablation_constructor_flags_match_reference_state_dict(flags):
    """transvae.py:36-38: the ablation switches change the module tree (no rope.inv_freq buffer, no dc_conv); the key list
    must be the one the reference builds (oracle.param_shapes is validated against it in validate_against_reference.py)."""
    m = transvae.TransVAE(config=MINI, latent_dim=32, **flags)
    want = O.init_state_dict(dict(MINI, **flags), seed=2)
    assert list(m.state_dict().keys()) == list(want.keys())
    assert all(tuple(v.shape) == tuple(want[k].shape) for k, v in m.state_dict().items())
    m.load_state_dict(want, strict=True)


def test_plain_ffn_ablation_is_rejected_like_the_reference():
    with pytest.raises(NotImplementedError, match="not functional in the reference"):
        transvae.TransVAE(config=MINI, latent_dim=32, use_conv_ffn=False)
