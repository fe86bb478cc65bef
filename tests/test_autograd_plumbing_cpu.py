"""Host logic of the block-level autograd Functions on CPU: the kernel layer (transvae.ops) is replaced by fp32 torch
emulations built on the addressing emulator (tests/emu.py), so what is tested is the plumbing of _autograd.ResBlockFn /
GroupNormSilu -- argument order, the GroupNorm-statistics hand-off from block to block (x_sums in, statistics of the
output out, marked non-differentiable), backward arity and the weight re-layout round trip -- against torch.autograd on
the reference formula  x + conv2(silu(GN2(conv1(silu(GN1(x))))))  (blocks.py:58-68).  No GPU, no .so compute calls."""
import pytest
import torch
import torch.nn.functional as F

from emu import emulate, emulate_wgrad
from transvae import _autograd as AG
from transvae import _taps as T
from transvae import ops


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def gn_sums(x_nhwc, groups):
    B, H, W, C = x_nhwc.shape
    v = x_nhwc.double().reshape(B, H * W, groups, C // groups)
    return torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1)


class FakeOps:
    """fp32 stand-ins with the signatures of the transvae.ops entry points the two Functions use."""

    def __init__(self):
        self.sums_given = 0

    def groupnorm_silu(self, x, gamma, beta, groups=32, eps=1e-5, silu=True, sums=None, return_sums=False):
        ref = gn_sums(x, groups)
        if sums is not None:
            self.sums_given += 1
            assert sums.dtype == torch.float64 and torch.allclose(sums, ref, rtol=1e-9, atol=1e-9)   # the right tensor's statistics
        y = F.group_norm(nchw(x.float()), groups, gamma.float(), beta.float(), eps)
        y = nhwc(F.silu(y) if silu else y)
        return (y, ref) if return_sums else y

    @staticmethod
    def gnb_part(x, dh, gamma, beta, groups, eps, silu):
        """Reduce pass of the GroupNorm backward: per (image, channel) (sum dy, sum dy * xhat), dy = dh * act'(gamma*xhat+beta)."""
        B, H, W, C = x.shape
        v = x.double().reshape(B, H * W, groups, C // groups)
        mean = v.mean(dim=(1, 3), keepdim=True)
        var = (v * v).mean(dim=(1, 3), keepdim=True) - mean * mean
        xhat = ((v - mean) / torch.sqrt(var + eps)).reshape(B, H * W, C)
        dy = dh.double().reshape(B, H * W, C)
        if silu:
            u = xhat * gamma.double() + beta.double()
            sg = torch.sigmoid(u)
            dy = dy * sg * (1 + u * (1 - sg))
        return torch.stack([dy.sum(1), (dy * xhat).sum(1)], dim=-1).float()

    def mtgemm(self, plan, a0, w, *, a1=None, out_shape=None, bias=None, residual=None, gn_groups=0, gn_bwd=None, **kw):
        assert not kw, kw
        out = emulate(plan, a0.float(), a1, w.float(), tuple(out_shape), bias=None if bias is None else bias.reshape(plan.num_phases, -1))
        if residual is not None:
            out = out + residual.float()
        if gn_groups:
            out._gn_sums = gn_sums(out, gn_groups)
        if gn_bwd is not None:                  # input-gradient GEMM that also leaves the GroupNorm-backward reduce pass
            gx, gs, gg, gb, ggroups, geps, gsilu = gn_bwd
            assert bias is None and residual is None and not gn_groups and gx.shape == out.shape
            assert torch.allclose(gs, gn_sums(gx, ggroups), rtol=1e-9, atol=1e-9)     # statistics of the right tensor
            out._gnb_part = self.gnb_part(gx, out, gg, gb, ggroups, geps, gsilu)
            self.parts_made = getattr(self, "parts_made", 0) + 1
        return out

    def mtgemm_wgrad(self, plan, a0, dz, n_total, a1=None, bias=False, dw_out=None, db_out=None):
        assert dw_out is None and db_out is None          # no trainer here: the plain autograd route
        dw = emulate_wgrad(plan, a0.float(), a1, dz.float(), n_total)
        if not bias:
            return dw
        return dw, dz.float().reshape(-1, n_total).sum(0, keepdim=True)

    def groupnorm_bwd(self, x, dh, sums, gamma, beta, add=None, groups=32, eps=1e-5, silu=True, part=None):
        assert torch.allclose(sums, gn_sums(x, groups), rtol=1e-9, atol=1e-9)
        if part is not None:                    # must be the reduce pass of exactly this (x, dh, gamma, beta)
            self.parts_used = getattr(self, "parts_used", 0) + 1
            assert torch.allclose(part, self.gnb_part(x, dh, gamma, beta, groups, eps, silu), rtol=1e-5, atol=1e-6)
        with torch.enable_grad():               # Function.backward runs with grad mode off
            xr = x.detach().float().clone().requires_grad_(True)
            g = gamma.detach().float().clone().requires_grad_(True)
            b = beta.detach().float().clone().requires_grad_(True)
            y = F.group_norm(nchw(xr), groups, g, b, eps)
            y = nhwc(F.silu(y) if silu else y)
            y.backward(dh.float())
        dx = xr.grad if add is None else xr.grad + add.float()
        return dx, g.grad, b.grad


@pytest.fixture
def fake(monkeypatch):
    f = FakeOps()
    for name in ("groupnorm_silu", "mtgemm", "mtgemm_wgrad", "groupnorm_bwd"):
        monkeypatch.setattr(ops, name, getattr(f, name))
    # weight re-layout through the torch route (the CUDA pack kernels are covered by the GPU tests)
    monkeypatch.setattr(AG, "_w_pack", lambda w: (T.pack_conv3x3(w.detach()), T.pack_conv3x3_dgrad(w.detach())))
    monkeypatch.setattr(AG, "_w_ungrad", lambda gp, w: gp.view(w.shape[0], 3, 3, w.shape[1]).permute(0, 3, 1, 2))
    monkeypatch.setattr(AG, "_f32", lambda t: t.detach().float().contiguous())
    return f


def _params(C, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, sc=1.0: torch.randn(*s, generator=g) * sc
    return dict(g1=1 + 0.2 * r(C), b1=0.1 * r(C), w1=r(C, C, 3, 3, sc=0.05), c1=0.1 * r(C),
                g2=1 + 0.2 * r(C), b2=0.1 * r(C), w2=r(C, C, 3, 3, sc=0.05), c2=0.1 * r(C))


def _ref_block(x_nchw, p):
    h = F.conv2d(F.silu(F.group_norm(x_nchw, 32, p["g1"], p["b1"], 1e-5)), p["w1"], p["c1"], padding=1)
    h = F.conv2d(F.silu(F.group_norm(h, 32, p["g2"], p["b2"], 1e-5)), p["w2"], p["c2"], padding=1)
    return x_nchw + h


def test_resblock_chain_hands_statistics_on_and_matches_autograd(fake):
    B, C, H, W = 2, 64, 6, 10
    x0 = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(0))
    pa, pb = _params(C, 1), _params(C, 2)
    leaves = {k: v.clone().requires_grad_(True) for k, v in list(("a." + k, v) for k, v in pa.items()) + list(("b." + k, v) for k, v in pb.items())}
    xin = nhwc(x0).requires_grad_(True)

    def run(x, x_sums, pre):
        q = lambda k: leaves[pre + k]
        return AG.ResBlockFn.apply(x, x_sums, q("g1"), q("b1"), q("w1"), q("c1"), q("g2"), q("b2"), q("w2"), q("c2"))

    y1, s1 = run(xin, None, "a.")
    assert not s1.requires_grad and s1.dtype == torch.float64 and tuple(s1.shape) == (B, 32, 2)
    assert torch.allclose(s1, gn_sums(y1.detach(), 32), rtol=1e-9, atol=1e-9)
    before = fake.sums_given
    y2, s2 = run(y1, s1, "b.")                      # the second block must use the first one's statistics for its GN1
    assert fake.sums_given == before + 2            # GN1 (handed on) and GN2 (conv1 epilogue)
    out = AG.GroupNormSilu.apply(y2, leaves["a.g1"] * 1.0, leaves["a.b1"] * 1.0, True, s2)   # decoder.norm_out hand-off
    dout = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    out.backward(dout)
    # both input-gradient GEMMs of both blocks left the reduce pass of the GroupNorm backward behind them
    assert fake.parts_made == 4 and fake.parts_used == 4

    # reference: torch.autograd on the NCHW formula
    ref_leaves = {k: v.detach().clone().requires_grad_(True) for k, v in leaves.items()}
    xr = x0.clone().requires_grad_(True)
    ra = {k[2:]: v for k, v in ref_leaves.items() if k.startswith("a.")}
    rb = {k[2:]: v for k, v in ref_leaves.items() if k.startswith("b.")}
    yr = _ref_block(_ref_block(xr, ra), rb)
    outr = F.silu(F.group_norm(yr, 32, ra["g1"] * 1.0, ra["b1"] * 1.0, 1e-5))
    outr.backward(nchw(dout))

    def close(a, b, what):
        err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
        assert err < 2e-4, (what, err)

    close(nchw(out.detach()), outr.detach(), "output")
    close(nchw(xin.grad), xr.grad, "dx")
    for k in leaves:
        close(leaves[k].grad, ref_leaves[k].grad, k)
