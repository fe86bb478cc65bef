"""Debug: run-to-run determinism, checkpointing, and bf16-autocast calibration of whole-model gradients."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("deepl-project_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import transvae, transvae_oracle as O
from util import build_model, load_golden, rel

blob, sd = load_golden("mini_tamed")
cfg = blob["cfg"]
m = build_model(cfg, sd).train()
x, eps = blob["x"].cuda(), blob["eps"].cuda()
loss_fn = transvae.TransVAELoss(lpips_weight=0.0, vf_weight=0.0, gan_weight=0.0)

gen = torch.Generator().manual_seed(5)
G = [torch.randn(blob["x"].shape, generator=gen).cuda() / blob["x"].numel(),
     torch.randn(blob["mu"].shape, generator=gen).cuda() / blob["mu"].numel(),
     torch.randn(blob["mu"].shape, generator=gen).cuda() / blob["mu"].numel()]

def objective(r, mu, lv):
    return (r.float() * G[0]).sum() + (mu.float() * G[1]).sum() + (lv.float() * G[2]).sum()

def grads():
    m.zero_grad()
    r, mu, lv = m(x, eps=eps)
    objective(r, mu, lv).backward()
    torch.cuda.synchronize()
    return {k: p.grad.detach().float().cpu().clone() for k, p in m.named_parameters()}

g0 = grads(); g1 = grads()
d = {k: rel(g1[k], g0[k]) for k in g0}
print("run-to-run worst:", sorted(d.items(), key=lambda kv: -kv[1])[:5])
m.enable_gradient_checkpointing()
g2 = grads()
d = {k: rel(g2[k], g0[k]) for k in g0}
print("ckpt vs plain worst:", sorted(d.items(), key=lambda kv: -kv[1])[:8])
first = [k for k, _ in m.named_parameters() if d[k] > 2e-2]
print("ckpt: params off by >2e-2:", len(first), first[:10], first[-5:])
l2 = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-30))
d01 = sorted(l2(g1[k], g0[k]) for k in g0); d02 = sorted(l2(g2[k], g0[k]) for k in g0)
print("l2 rel: run-to-run median %.4f max %.4f | ckpt-vs-plain median %.4f max %.4f" % (d01[len(d01)//2], d01[-1], d02[len(d02)//2], d02[-1]))
# single block with / without checkpoint
import torch.utils.checkpoint as cp
for name, blk, shape in (("res", m.encoder.stages[0][0], (2, 32, 32, 64)), ("tvb", m.encoder.stages[3][0], (2, 8, 8, 128)),
                         ("down", m.encoder.downsamples[2], (2, 16, 16, 64))):
    xb = (torch.randn(shape, generator=torch.Generator().manual_seed(3)) * 2).to(torch.bfloat16).cuda()
    dout = None
    res = []
    for use in (False, True, False):
        blk.zero_grad()
        xi = xb.clone().requires_grad_(True)
        out = cp.checkpoint(blk.forward_nhwc, xi, use_reentrant=False) if use else blk.forward_nhwc(xi)
        if dout is None:
            dout = torch.randn(out.shape, generator=torch.Generator().manual_seed(4)).to(torch.bfloat16).cuda()
        out.backward(dout)
        res.append((xi.grad.float().clone(), {k: p.grad.float().clone() for k, p in blk.named_parameters()}))
    print(name, "dx ckpt-vs-plain", l2(res[1][0], res[0][0]), "plain-vs-plain", l2(res[2][0], res[0][0]),
          "worst param ckpt", max(l2(res[1][1][k], res[0][1][k]) for k in res[0][1]),
          "worst param plain", max(l2(res[2][1][k], res[0][1][k]) for k in res[0][1]))

# calibration: oracle fp32 vs oracle under bf16 autocast on the GPU
def oracle_grads(autocast):
    sdg = {k: v.cuda().clone().requires_grad_("inv_freq" not in k) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        rec, mu, lv, _ = O.forward(sdg, cfg, x, eps, patched=True)
    objective(rec, mu, lv).backward()
    return {k: v.grad.detach().float().cpu() for k, v in sdg.items() if v.requires_grad}

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ref = oracle_grads(False)
ac = oracle_grads(True)
rows = []
for k in g0:
    r = ref[k].flatten(); n = float(r.norm())
    if n < 1e-12: continue
    e_ours = float((g0[k].flatten() - r).norm()) / n
    e_ac = float((ac[k].flatten() - r).norm()) / n
    rows.append((e_ours / max(e_ac, 1e-6), e_ours, e_ac, k))
rows.sort(reverse=True)
print("worst ratio ours/autocast (l2 rel err):")
for r in rows[:15]: print("  ratio %.2f ours %.4f autocast %.4f %s" % r)
print("median ours %.4f median autocast %.4f" % (sorted(r[1] for r in rows)[len(rows)//2], sorted(r[2] for r in rows)[len(rows)//2]))
