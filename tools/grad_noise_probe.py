"""Run-to-run reproducibility probe of the TRAINING step: the same forward + backward twice on the same input / eps,
recording the output of every kernel launch wrapper in call order, to find which kernel first introduces a difference.
    python tools/grad_noise_probe.py [mini_tamed|mini_tamed_128|large] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("deepl-project_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import transvae
from transvae import ops

which = sys.argv[1] if len(sys.argv) > 1 else "mini_tamed"
if which.startswith("mini"):
    from util import build_model, load_golden
    blob, sd = load_golden(which)
    m = build_model(blob["cfg"], sd).train()
    x, eps = blob["x"].cuda(), blob["eps"].cuda()
else:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    torch.manual_seed(0)
    with torch.device("cuda"):
        m = transvae.TransVAE(variant="large", compression_ratio=16, latent_dim=32).train()
    x = torch.rand(B, 3, 256, 256, generator=torch.Generator().manual_seed(1)).cuda()
    eps = torch.randn(B, 32, 16, 16, generator=torch.Generator().manual_seed(2)).cuda()

NAMES = ["mtgemm", "attn_fwd", "groupnorm_silu", "groupnorm_bwd", "token_norm_fwd", "token_norm_bwd", "attn_bwd", "act_fwd",
         "mtgemm_wgrad", "bias_act_bwd", "conv_in", "conv_in_wgrad", "loss_bwd", "latent_bwd", "reparam", "nchw_to_nhwc",
         "nhwc_to_nchw", "wgrad_unpack", "weight_pack", "loss_sums", "im2col_in"]
REC = None
orig = {n: getattr(ops, n) for n in NAMES if hasattr(ops, n)}


def flat(o):
    if isinstance(o, torch.Tensor):
        return [o]
    if isinstance(o, (tuple, list)):
        return [t for e in o for t in flat(e)]
    return []


def wrap(name, fn):
    def w(*a, **k):
        out = fn(*a, **k)
        if REC is not None:
            tag = name
            if name in ("mtgemm", "mtgemm_wgrad"):
                tag += ":" + a[0].name
            ts = flat(out)
            extra = getattr(ts[0], "_gn_sums", None) if ts else None
            REC.append((tag, [t.detach().float().clone() for t in ts] + ([extra.float().clone()] if extra is not None else [])))
        return out
    return w


for n, f in orig.items():
    setattr(ops, n, wrap(n, f))
import transvae._autograd as AG  # noqa: E402  (uses ops.<name> at call time)


def run():
    global REC
    REC = []
    m.zero_grad()
    loss_fn = transvae.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)
    rec, mu, lv = m(x, eps=eps)
    loss_fn(rec, x, mu, lv)["total"].backward()
    torch.cuda.synchronize()
    r, REC = REC, None
    grads = {k: p.grad.detach().float().clone() for k, p in m.named_parameters()}
    return r, grads


(a, ga), (b, gb) = run(), run()
assert len(a) == len(b)
first, shown = None, 0
for i, ((na, ta), (nb, tb)) in enumerate(zip(a, b)):
    for j, (u, v) in enumerate(zip(ta, tb)):
        d = u - v
        nz = int((d != 0).sum())
        if nz:
            if first is None:
                first = (i, na, j)
            if shown < 40:
                shown += 1
                print(f"#{i:4d} {na:44s} out{j} differing {nz}/{d.numel()}  l2 {float(d.norm() / u.norm().clamp_min(1e-30)):.2e}  "
                      f"max {float(d.abs().max() / u.abs().max().clamp_min(1e-30)):.2e}")
print("launch wrappers recorded:", len(a), " first differing:", first)
errs = sorted((float((ga[k] - gb[k]).norm() / ga[k].norm().clamp_min(1e-30)), k) for k in ga)
print("parameter gradients: median l2 %.3e  max %.3e (%s); bit-identical tensors %d / %d" %
      (errs[len(errs) // 2][0], errs[-1][0], errs[-1][1], sum(1 for e, _ in errs if e == 0.0), len(errs)))
