"""CPU oracle for the TransVAE hot path -- TEST INFRASTRUCTURE ONLY.

This file is a self-contained fp32 restatement (plain PyTorch ops, functional
style over a flat ``state_dict``) of the reference's TransVAE encoder/decoder
forward, the reparameterisation and the L1+KL loss.  It is the *checker* for
the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline / ``--impl reference`` legs may import it.  Nothing under
``deepl-project_b200/`` imports it, and the product never falls back to it.

Parity status: the reference ships no golden vectors or numeric tests for this
path ("parity unpinned" by the reference's own tests).  The
oracle is therefore pinned against the reference itself:
``oracle/validate_against_reference.py`` imports the unmodified reference from
/root/reference in the authoring container, loads the same seeded weights and
checks that every function below reproduces it bit-for-bit on CPU;
``oracle/make_golden.py`` stores reference outputs under ``tests/golden/``.

The arithmetic (conv2d, linear, group_norm, layer_norm, SDPA, gelu, silu,
pixel_(un)shuffle) lives in third-party PyTorch (un-pinned ``torch>=2.0.0`` in
the reference's requirements; 2.11.0+cu128 in this image); the reference's own
arithmetic is RMSNorm, RoPE2D, the reparameterisation and the loss formulas.

Reference files restated (T = /root/reference/transvae-implementation):
  T/transvae/models/transvae.py   encode :170, reparameterize :186, decode :201,
                                  forward :213, _get_variant_config :107
  T/transvae/models/encoder.py    forward :101
  T/transvae/models/decoder.py    forward :102
  T/transvae/modules/blocks.py    ResBlock :48, TransVAEBlock :135, RMSNorm :168
  T/transvae/modules/attention.py FlashAttentionWithRoPE :55, RoPE2D :132
  T/transvae/modules/conv.py      ConvFFN :69
  T/transvae/modules/upsample.py  Downsample :44, Upsample :105
  T/transvae/losses/vae_loss.py   L1 :83, KL :94-95
  T/transvae-implementation_patched/transvae/models/transvae.py  :186-196, :244-245
  T/transvae-implementation_patched/transvae/losses/vae_loss.py  :80-104
"""
from __future__ import annotations

import hashlib
import math
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

# ----------------------------------------------------------------------------
# configuration table (T/transvae/models/transvae.py:107-153)
# ----------------------------------------------------------------------------
VARIANTS = {
    "tiny_f16d32": dict(depths=[3, 3, 3, 3, 3], base_dims=[128, 128, 256, 256, 512]),
    "base_f16d32": dict(depths=[3, 3, 3, 3, 3], base_dims=[128, 128, 256, 512, 1024]),
    "large_f16d32": dict(depths=[3, 3, 3, 4, 6], base_dims=[192, 192, 384, 768, 1536]),
    "huge_f16d32": dict(depths=[3, 3, 4, 6, 8], base_dims=[256, 256, 512, 1024, 2048]),
    "giant_f16d32": dict(depths=[3, 3, 4, 8, 10], base_dims=[320, 320, 640, 1280, 2560]),
    "large_f8d16": dict(depths=[3, 3, 6, 8], base_dims=[192, 384, 768, 1536]),
}


def variant_config(variant: str = "large", f: int = 16, d: int = 32) -> dict:
    key = f"{variant}_f{f}d{d}"
    if key not in VARIANTS:
        raise ValueError(f"Unknown variant: {variant} with f{f}d{d}")
    cfg = dict(VARIANTS[key])
    cfg.update(mlp_ratio=1.0, head_dim=64, latent_dim=d, input_channels=3)
    return cfg


def normalise_config(cfg: dict) -> dict:
    out = dict(cfg)
    out.setdefault("mlp_ratio", 1.0)
    out.setdefault("head_dim", 64)
    out.setdefault("latent_dim", 32)
    out.setdefault("input_channels", 3)
    # ablation switches of the reference constructor (transvae.py:36-38); defaults = the shipped configs
    out.setdefault("use_rope", True)
    out.setdefault("use_conv_ffn", True)
    out.setdefault("use_dc_path", True)
    return out


# ----------------------------------------------------------------------------
# parameter inventory: key -> shape, in the reference's state_dict naming
# ----------------------------------------------------------------------------
def _block_shapes(prefix: str, dim: int, mlp_ratio: float, head_dim: int, use_rope: bool = True,
                  use_conv_ffn: bool = True) -> List[Tuple[str, Tuple[int, ...], str]]:
    hid = int(dim * mlp_ratio * 4)
    mid = int(dim * mlp_ratio)
    p = prefix
    out = _block_shapes_full(p, dim, hid, mid, head_dim)
    if not use_rope:                       # attention.py:50-53: no RoPE2D submodule, hence no inv_freq buffer
        out = [e for e in out if not e[0].endswith("attn.rope.inv_freq")]
    if not use_conv_ffn:                   # blocks.py:124-133: nn.Sequential(Linear, GELU, Dropout, Linear, Dropout)
        out = [e for e in out if ".ffn." not in e[0]]
        h2 = int(dim * mlp_ratio)
        out += [(p + "ffn.0.weight", (h2, dim), "linear"), (p + "ffn.0.bias", (h2,), "bias"),
                (p + "ffn.3.weight", (dim, h2), "linear"), (p + "ffn.3.bias", (dim,), "bias")]
    return out


def _block_shapes_full(p: str, dim: int, hid: int, mid: int, head_dim: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    return [
        (p + "norm1.weight", (dim,), "norm_w"),
        (p + "attn.norm_q.weight", (dim,), "norm_w"), (p + "attn.norm_q.bias", (dim,), "bias"),
        (p + "attn.norm_k.weight", (dim,), "norm_w"), (p + "attn.norm_k.bias", (dim,), "bias"),
        (p + "attn.norm_v.weight", (dim,), "norm_w"), (p + "attn.norm_v.bias", (dim,), "bias"),
        (p + "attn.to_q.weight", (dim, dim), "linear"),
        (p + "attn.to_k.weight", (dim, dim), "linear"),
        (p + "attn.to_v.weight", (dim, dim), "linear"),
        (p + "attn.proj.weight", (dim, dim), "linear"), (p + "attn.proj.bias", (dim,), "bias"),
        (p + "attn.rope.inv_freq", (head_dim // 4,), "inv_freq"),
        (p + "norm2.weight", (dim,), "norm_w"),
        (p + "ffn.proj_in.weight", (hid, dim), "linear"), (p + "ffn.proj_in.bias", (hid,), "bias"),
        (p + "ffn.conv.0.weight", (mid, hid, 1, 1), "conv"), (p + "ffn.conv.0.bias", (mid,), "bias"),
        (p + "ffn.conv.2.weight", (mid, mid, 3, 3), "conv"), (p + "ffn.conv.2.bias", (mid,), "bias"),
        (p + "ffn.conv.4.weight", (hid, mid, 1, 1), "conv"), (p + "ffn.conv.4.bias", (hid,), "bias"),
        (p + "ffn.proj_out.weight", (dim, hid), "linear"), (p + "ffn.proj_out.bias", (dim,), "bias"),
    ]


def _resblock_shapes(prefix: str, cin: int, cout: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    p = prefix
    out = [
        (p + "norm1.weight", (cin,), "norm_w"), (p + "norm1.bias", (cin,), "bias"),
        (p + "conv1.weight", (cout, cin, 3, 3), "conv"), (p + "conv1.bias", (cout,), "bias"),
        (p + "norm2.weight", (cout,), "norm_w"), (p + "norm2.bias", (cout,), "bias"),
        (p + "conv2.weight", (cout, cout, 3, 3), "conv"), (p + "conv2.bias", (cout,), "bias"),
    ]
    if cin != cout:
        out += [(p + "shortcut.weight", (cout, cin, 1, 1), "conv"), (p + "shortcut.bias", (cout,), "bias")]
    return out


def param_shapes(cfg: dict) -> List[Tuple[str, Tuple[int, ...], str]]:
    """All state_dict entries (name, shape, kind) in the reference's order."""
    cfg = normalise_config(cfg)
    depths, dims = cfg["depths"], cfg["base_dims"]
    mr, hd, d, cin = cfg["mlp_ratio"], cfg["head_dim"], cfg["latent_dim"], cfg["input_channels"]
    rope, cffn, dc = cfg["use_rope"], cfg["use_conv_ffn"], cfg["use_dc_path"]
    n = len(depths)
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    out += [("encoder.conv_in.weight", (dims[0], cin, 3, 3), "conv"), ("encoder.conv_in.bias", (dims[0],), "bias")]
    for i in range(n):
        for j in range(depths[i]):
            p = f"encoder.stages.{i}.{j}."
            out += _resblock_shapes(p, dims[i], dims[i]) if i < 2 else _block_shapes(p, dims[i], mr, hd, rope, cffn)
    for i in range(n - 1):
        p = f"encoder.downsamples.{i}."
        out += [
            (p + "main_path.0.weight", (dims[i], dims[i], 3, 3), "conv"), (p + "main_path.0.bias", (dims[i],), "bias"),
            (p + "main_path.2.weight", (dims[i + 1], dims[i], 3, 3), "conv"), (p + "main_path.2.bias", (dims[i + 1],), "bias"),
        ]
        if dc:
            out += [(p + "dc_conv.weight", (dims[i + 1], dims[i] * 4, 1, 1), "conv"), (p + "dc_conv.bias", (dims[i + 1],), "bias")]
    out += [("conv_mu.weight", (d, dims[-1], 3, 3), "conv"), ("conv_mu.bias", (d,), "bias"),
            ("conv_logvar.weight", (d, dims[-1], 3, 3), "conv"), ("conv_logvar.bias", (d,), "bias")]
    rd, rdep = dims[::-1], depths[::-1]
    out += [("decoder.conv_in.weight", (rd[0], d, 3, 3), "conv"), ("decoder.conv_in.bias", (rd[0],), "bias")]
    for i in range(n):
        for j in range(rdep[i]):
            p = f"decoder.stages.{i}.{j}."
            out += _block_shapes(p, rd[i], mr, hd, rope, cffn) if i < n - 2 else _resblock_shapes(p, rd[i], rd[i])
    for i in range(n - 1):
        p = f"decoder.upsamples.{i}."
        out += [
            (p + "main_path.1.weight", (rd[i + 1], rd[i], 3, 3), "conv"), (p + "main_path.1.bias", (rd[i + 1],), "bias"),
            (p + "main_path.3.weight", (rd[i + 1], rd[i + 1], 3, 3), "conv"), (p + "main_path.3.bias", (rd[i + 1],), "bias"),
        ]
        if dc:
            out += [(p + "dc_conv.weight", (rd[i + 1] * 4, rd[i], 1, 1), "conv"), (p + "dc_conv.bias", (rd[i + 1] * 4,), "bias")]
    out += [("decoder.norm_out.weight", (rd[-1],), "norm_w"), ("decoder.norm_out.bias", (rd[-1],), "bias"),
            ("decoder.conv_out.weight", (cin, rd[-1], 3, 3), "conv"), ("decoder.conv_out.bias", (cin,), "bias")]
    return out


def _key_seed(seed: int, key: str) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def init_state_dict(cfg: dict, seed: int = 0, mode: str = "reference") -> StateDict:
    """Deterministic fp32 weights keyed like the reference's ``state_dict``.

    Each tensor is drawn from its own ``torch.Generator`` seeded from
    ``sha256(seed:key)`` so the values do not depend on construction order.
    ``mode='reference'`` follows the distributions of
    ``TransVAE._initialize_weights`` (transvae.py:155-168): Kaiming-normal
    (fan_out, relu) convs, trunc-normal(0.02) linears, zero biases, unit norms.
    ``mode='tamed'`` additionally randomises biases and norm
    scales and shrinks the heads / residual-branch output layers so that clamps do
    not saturate and every parameter influences the output measurably.
    """
    sd: StateDict = {}
    for key, shape, kind in param_shapes(cfg):
        g = torch.Generator().manual_seed(_key_seed(seed, key))
        if kind == "inv_freq":
            dpa = shape[0] * 2
            sd[key] = 1.0 / (10000 ** (torch.arange(0, dpa, 2).float() / dpa))
            continue
        if kind == "conv":
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
        elif kind == "linear":
            t = (torch.randn(shape, generator=g) * 0.02).clamp_(-2.0, 2.0)   # trunc_normal_(std=.02, a=-2, b=2)
        elif kind == "norm_w":
            t = torch.ones(shape)
            if mode == "tamed":
                t = t + 0.1 * torch.randn(shape, generator=g)
        elif kind == "bias":
            t = torch.zeros(shape)
            if mode == "tamed":
                t = 0.05 * torch.randn(shape, generator=g)
        else:  # pragma: no cover
            raise AssertionError(kind)
        if mode == "tamed":
            if key.startswith(("conv_mu.", "conv_logvar.")):
                t = t * 0.02
            elif kind == "linear":
                t = t * 8.0           # make attention / FFN branches comparable to the stream
            elif kind == "conv" and (".downsamples." in key or ".upsamples." in key):
                t = t * 0.6
        sd[key] = t.contiguous()
    return sd


# ----------------------------------------------------------------------------
# modules (functional)
# ----------------------------------------------------------------------------
def rmsnorm(x: Tensor, w: Tensor, eps: float = 1e-6) -> Tensor:
    """blocks.py:168-201 (4-D branch): reduce over channels per pixel."""
    B, C, H, W = x.shape
    xf = x.view(B, C, -1)
    rms = torch.sqrt(torch.mean(xf ** 2, dim=1, keepdim=True) + eps)
    y = xf / rms
    y = y * w.view(1, -1, 1)
    return y.view(B, C, H, W)


def rope2d(x: Tensor, H: int, W: int, inv_freq: Tensor) -> Tensor:
    """attention.py:132-199.  x: [B, heads, N, hd].  NOT a rotation:
    even outputs use the angle of slot 2i, odd outputs the angle of slot 2i+1."""
    B, nh, N, hd = x.shape
    yp = torch.arange(H, device=x.device, dtype=x.dtype)
    xp = torch.arange(W, device=x.device, dtype=x.dtype)
    yg, xg = torch.meshgrid(yp, xp, indexing="ij")
    yf = torch.outer(yg.flatten(), inv_freq)
    xf = torch.outer(xg.flatten(), inv_freq)
    ang = torch.cat([yf, yf, xf, xf], dim=-1)            # [N, hd]
    cs, sn = ang.cos().view(1, 1, N, hd // 2, 2), ang.sin().view(1, 1, N, hd // 2, 2)
    xr = x.view(B, nh, N, hd // 2, 2)
    a, b = xr[..., 0], xr[..., 1]
    o1 = a * cs[..., 0] - b * sn[..., 0]
    o2 = a * sn[..., 1] + b * cs[..., 1]
    return torch.stack([o1, o2], dim=-1).view(B, nh, N, hd)


def rope_angles(H: int, W: int, inv_freq: Tensor) -> Tensor:
    """Closed form of the angle table theta[n, j], fp32, [H*W, hd]."""
    r = torch.arange(H, dtype=torch.float32).repeat_interleave(W)
    c = torch.arange(W, dtype=torch.float32).repeat(H)
    f = inv_freq.float()
    return torch.cat([torch.outer(r, f), torch.outer(r, f), torch.outer(c, f), torch.outer(c, f)], dim=-1)


def attention(sd: StateDict, p: str, x: Tensor, head_dim: int, use_rope: bool = True,
              trace: Optional[dict] = None) -> Tensor:
    """attention.py:55-104.  x: [B, C, H, W] (already RMS-normalised by the caller)."""
    B, C, H, W = x.shape
    nh = C // head_dim
    t = x.flatten(2).transpose(1, 2)
    q = F.linear(F.layer_norm(t, (C,), sd[p + "norm_q.weight"], sd[p + "norm_q.bias"], 1e-5), sd[p + "to_q.weight"])
    k = F.linear(F.layer_norm(t, (C,), sd[p + "norm_k.weight"], sd[p + "norm_k.bias"], 1e-5), sd[p + "to_k.weight"])
    v = F.linear(F.layer_norm(t, (C,), sd[p + "norm_v.weight"], sd[p + "norm_v.bias"], 1e-5), sd[p + "to_v.weight"])
    q = q.view(B, H * W, nh, head_dim).transpose(1, 2)
    k = k.view(B, H * W, nh, head_dim).transpose(1, 2)
    v = v.view(B, H * W, nh, head_dim).transpose(1, 2)
    if use_rope and (p + "rope.inv_freq") in sd:      # use_rope=False builds no RoPE2D module (attention.py:50-53)
        q = rope2d(q, H, W, sd[p + "rope.inv_freq"])
        k = rope2d(k, H, W, sd[p + "rope.inv_freq"])
    if trace is not None:
        trace[p + "q"], trace[p + "k"], trace[p + "v"] = q, k, v
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, scale=head_dim ** -0.5)
    o = o.transpose(1, 2).reshape(B, H * W, C)
    if trace is not None:
        trace[p + "o"] = o
    o = F.linear(o, sd[p + "proj.weight"], sd[p + "proj.bias"])
    return o.transpose(1, 2).reshape(B, C, H, W)


def conv_ffn(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """conv.py:69-105, conv_type='full' (the only variant the shipped configs reach)."""
    B, C, H, W = x.shape
    t = x.flatten(2).transpose(1, 2)
    u = F.gelu(F.linear(t, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"]))
    s = u.transpose(1, 2).reshape(B, -1, H, W)
    if (p + "conv.weight") in sd:                    # conv_type='depthwise' (conv.py:42-50): one grouped 3x3, no activation
        c = F.conv2d(s, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1, groups=s.shape[1])
        t = (s + c).flatten(2).transpose(1, 2)
        o = F.linear(t, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
        return o.transpose(1, 2).reshape(B, C, H, W)
    c = F.conv2d(s, sd[p + "conv.0.weight"], sd[p + "conv.0.bias"])
    c = F.gelu(c)
    c = F.conv2d(c, sd[p + "conv.2.weight"], sd[p + "conv.2.bias"], padding=1)
    c = F.gelu(c)
    c = F.conv2d(c, sd[p + "conv.4.weight"], sd[p + "conv.4.bias"])
    s = s + c
    t = s.flatten(2).transpose(1, 2)
    o = F.linear(t, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return o.transpose(1, 2).reshape(B, C, H, W)


def plain_ffn(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """blocks.py:124-133 (use_conv_ffn=False): Linear -> GELU -> Linear applied to the LAST dimension of the NCHW tensor
    exactly as the reference's nn.Sequential does (blocks.py:149) -- i.e. along W, which only type-checks when W == dim."""
    h = F.gelu(F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"]))
    return F.linear(h, sd[p + "3.weight"], sd[p + "3.bias"])


def transvae_block(sd: StateDict, p: str, x: Tensor, head_dim: int, trace: Optional[dict] = None) -> Tensor:
    """blocks.py:135-151."""
    a = attention(sd, p + "attn.", rmsnorm(x, sd[p + "norm1.weight"]), head_dim, trace=trace)
    x = x + a
    if (p + "ffn.proj_in.weight") in sd:
        f = conv_ffn(sd, p + "ffn.", rmsnorm(x, sd[p + "norm2.weight"]))
    else:
        f = plain_ffn(sd, p + "ffn.", rmsnorm(x, sd[p + "norm2.weight"]))
    if trace is not None:
        trace[p + "attn_branch"], trace[p + "ffn_branch"] = a, f
    return x + f


def resblock(sd: StateDict, p: str, x: Tensor, trace: Optional[dict] = None) -> Tensor:
    """blocks.py:48-68."""
    h = F.group_norm(x, 32, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    h = F.silu(h)
    h = F.conv2d(h, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    h = F.group_norm(h, 32, sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    h = F.silu(h)
    h = F.conv2d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if trace is not None:
        trace[p + "branch"] = h
    if (p + "shortcut.weight") in sd:             # blocks.py:40-46: 1x1, or 3x3 (padding 1) with use_conv_shortcut
        x = F.conv2d(x, sd[p + "shortcut.weight"], sd[p + "shortcut.bias"], padding=sd[p + "shortcut.weight"].shape[-1] // 2)
    return h + x


def downsample(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """upsample.py:44-66."""
    m = F.conv2d(x, sd[p + "main_path.0.weight"], sd[p + "main_path.0.bias"], padding=1)
    m = F.silu(m)
    m = F.conv2d(m, sd[p + "main_path.2.weight"], sd[p + "main_path.2.bias"], stride=2, padding=1)
    if (p + "dc_conv.weight") in sd:
        dc = F.conv2d(F.pixel_unshuffle(x, 2), sd[p + "dc_conv.weight"], sd[p + "dc_conv.bias"])
        m = m + dc
    return m


def upsample(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """upsample.py:105-128."""
    m = F.interpolate(x, scale_factor=2.0, mode="nearest")
    m = F.conv2d(m, sd[p + "main_path.1.weight"], sd[p + "main_path.1.bias"], padding=1)
    m = F.silu(m)
    m = F.conv2d(m, sd[p + "main_path.3.weight"], sd[p + "main_path.3.bias"], padding=1)
    if (p + "dc_conv.weight") in sd:
        dc = F.pixel_shuffle(F.conv2d(x, sd[p + "dc_conv.weight"], sd[p + "dc_conv.bias"]), 2)
        m = m + dc
    return m


# ----------------------------------------------------------------------------
# model
# ----------------------------------------------------------------------------
def encoder_forward(sd: StateDict, cfg: dict, x: Tensor, trace: Optional[dict] = None) -> Tensor:
    """encoder.py:101-126."""
    cfg = normalise_config(cfg)
    depths = cfg["depths"]
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    if trace is not None:
        trace["encoder.conv_in"] = h
    for i, dep in enumerate(depths):
        for j in range(dep):
            p = f"encoder.stages.{i}.{j}."
            h = resblock(sd, p, h, trace) if i < 2 else transvae_block(sd, p, h, cfg["head_dim"], trace)
            if trace is not None:
                trace[p[:-1]] = h
        if i < len(depths) - 1:
            h = downsample(sd, f"encoder.downsamples.{i}.", h)
            if trace is not None:
                trace[f"encoder.downsamples.{i}"] = h
    return h


def decoder_forward(sd: StateDict, cfg: dict, z: Tensor, trace: Optional[dict] = None) -> Tensor:
    """decoder.py:102-132."""
    cfg = normalise_config(cfg)
    depths = cfg["depths"][::-1]
    n = len(depths)
    h = F.conv2d(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    if trace is not None:
        trace["decoder.conv_in"] = h
    for i, dep in enumerate(depths):
        for j in range(dep):
            p = f"decoder.stages.{i}.{j}."
            h = transvae_block(sd, p, h, cfg["head_dim"], trace) if i < n - 2 else resblock(sd, p, h, trace)
            if trace is not None:
                trace[p[:-1]] = h
        if i < n - 1:
            h = upsample(sd, f"decoder.upsamples.{i}.", h)
            if trace is not None:
                trace[f"decoder.upsamples.{i}"] = h
    h = F.group_norm(h, 32, sd["decoder.norm_out.weight"], sd["decoder.norm_out.bias"], 1e-5)
    h = F.silu(h)
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


def encode(sd: StateDict, cfg: dict, x: Tensor, trace: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """transvae.py:170-184."""
    h = encoder_forward(sd, cfg, x, trace)
    mu = F.conv2d(h, sd["conv_mu.weight"], sd["conv_mu.bias"], padding=1)
    logvar = F.conv2d(h, sd["conv_logvar.weight"], sd["conv_logvar.bias"], padding=1)
    return mu, logvar


def decode(sd: StateDict, cfg: dict, z: Tensor, trace: Optional[dict] = None) -> Tensor:
    """transvae.py:201-211."""
    return decoder_forward(sd, cfg, z, trace)


def reparameterize(mu: Tensor, logvar: Tensor, eps: Tensor, patched: bool = True) -> Tensor:
    """transvae.py:186-199 / patched :186-196.  ``eps`` is passed in (the reference
    draws it with randn_like; the caller draws it with the same generator)."""
    if patched:
        std = torch.exp(0.5 * logvar.float().clamp(-30.0, 20.0))
        return (mu.float() + eps * std).to(mu.dtype)
    return mu + eps * torch.exp(0.5 * logvar)


def forward(sd: StateDict, cfg: dict, x: Tensor, eps: Tensor, patched: bool = True,
            trace: Optional[dict] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """transvae.py:213-242 (+ patched clamps :244-245).  Returns (recon, mu, logvar, z)."""
    mu, logvar = encode(sd, cfg, x, trace)
    if patched:
        mu = mu.clamp(-50, 50)
        logvar = logvar.clamp(-30, 20)
    z = reparameterize(mu, logvar, eps, patched)
    return decode(sd, cfg, z, trace), mu, logvar, z


def loss_l1_kl(recon: Tensor, target: Tensor, mu: Tensor, logvar: Tensor, l1_weight: float = 1.0,
               kl_weight: float = 1e-8, patched: bool = True,
               logvar_clip: Tuple[float, float] = (-30.0, 20.0)) -> Dict[str, Tensor]:
    """L1 + KL of TransVAELoss.forward.  main: vae_loss.py:83, :94-95 (KL summed and
    divided by B*H*W).  patched: vae_loss.py:80-104 (sigmoid on recon, fp32 clamped KL,
    mean over all elements)."""
    if patched:
        l1 = F.l1_loss(recon.sigmoid(), target)
        lv = logvar.float().clamp(logvar_clip[0], logvar_clip[1])
        kl = (-0.5 * (1.0 + lv - mu.float().pow(2) - lv.exp())).mean()
    else:
        l1 = F.l1_loss(recon, target)
        kl = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
        kl = kl / (mu.shape[0] * mu.shape[2] * mu.shape[3])
    out = {"l1": l1 * l1_weight, "kl": kl * kl_weight}
    out["total"] = out["l1"] + out["kl"]
    return out


def psnr(a: Tensor, b: Tensor, max_val: float = 1.0) -> float:
    """patched/evaluate_transvae.py:47-53."""
    mse = F.mse_loss(a.float(), b.float())
    if mse == 0:
        return float("inf")
    return float(20 * torch.log10(torch.tensor(max_val) / torch.sqrt(mse)))


def ssim(img1: Tensor, img2: Tensor, window_size: int = 11, size_average: bool = True):
    """patched/evaluate_transvae.py:56-77: box-filter SSIM (avg_pool2d, zero padding counted in the mean)."""
    img1, img2 = img1.float(), img2.float()
    pad = window_size // 2
    pool = lambda t: F.avg_pool2d(t, window_size, stride=1, padding=pad)
    mu1, mu2 = pool(img1), pool(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = pool(img1 * img1) - mu1_sq
    sigma2_sq = pool(img2 * img2) - mu2_sq
    sigma12 = pool(img1 * img2) - mu1_mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + c1) * (2 * sigma12 + c2)) / ((mu1_sq + mu2_sq + c1) * (sigma1_sq + sigma2_sq + c2))
    if size_average:
        return float(ssim_map.mean())
    return ssim_map.mean(1).mean(1).mean(1)


def max_rel_err(ours: Tensor, ref: Tensor) -> float:
    """Parity metric of SURVEY 8c: max|ours-ref| / max|ref|."""
    ref = ref.float()
    return float((ours.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def count_params(cfg: dict) -> Dict[str, int]:
    tot = enc = dec = 0
    for key, shape, kind in param_shapes(cfg):
        if kind == "inv_freq":
            continue
        n = math.prod(shape)
        tot += n
        if key.startswith("encoder."):
            enc += n
        elif key.startswith("decoder."):
            dec += n
    return {"encoder": enc, "decoder": dec, "total": tot}


def forward_flops_per_image(cfg: dict, res: int) -> float:
    """Algorithmic forward FLOPs per image (2*M*N*K per conv/linear, 4*S^2*C per SDPA),
    the counting rule of SURVEY section 6."""
    cfg = normalise_config(cfg)
    depths, dims, mr = cfg["depths"], cfg["base_dims"], cfg["mlp_ratio"]
    d, cin, n = cfg["latent_dim"], cfg["input_channels"], len(cfg["depths"])

    def conv(hw, ci, co, k):
        return 2.0 * hw * co * ci * k * k

    def block(hw, c):
        hid, mid = int(c * mr * 4), int(c * mr)
        return (4 * conv(hw, c, c, 1) + 4.0 * hw * hw * c + conv(hw, c, hid, 1) + conv(hw, hid, mid, 1)
                + conv(hw, mid, mid, 3) + conv(hw, mid, hid, 1) + conv(hw, hid, c, 1))

    fl = 0.0
    r = res
    fl += conv(r * r, cin, dims[0], 3)
    for i in range(n):
        hw = r * r
        fl += depths[i] * (2 * conv(hw, dims[i], dims[i], 3) if i < 2 else block(hw, dims[i]))
        if i < n - 1:
            fl += conv(hw, dims[i], dims[i], 3) + conv(hw // 4, dims[i], dims[i + 1], 3) + conv(hw // 4, 4 * dims[i], dims[i + 1], 1)
            r //= 2
    fl += 2 * conv(r * r, dims[-1], d, 3)
    rd, rdep = dims[::-1], depths[::-1]
    fl += conv(r * r, d, rd[0], 3)
    for i in range(n):
        hw = r * r
        fl += rdep[i] * (block(hw, rd[i]) if i < n - 2 else 2 * conv(hw, rd[i], rd[i], 3))
        if i < n - 1:
            fl += conv(hw * 4, rd[i], rd[i + 1], 3) + conv(hw * 4, rd[i + 1], rd[i + 1], 3) + conv(hw, rd[i], 4 * rd[i + 1], 1)
            r *= 2
    fl += conv(r * r, rd[-1], cin, 3)
    return fl
