"""Run-to-run reproducibility probe: the same encode (+ decode) twice on the same input, per-layer l2 / max difference
(encoder / decoder `trace` hooks), to find which kernel introduces non-determinism.
    python tools/noise_probe.py [mini|mini_tamed_128|large] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepl-project_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import transvae

which = sys.argv[1] if len(sys.argv) > 1 else "mini"
if which.startswith("mini"):
    from util import build_model, load_golden
    blob, sd = load_golden(which if which != "mini" else "mini_tamed")
    m = build_model(blob["cfg"], sd)
    x = blob["x"].cuda()
else:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    torch.manual_seed(0)
    with torch.device("cuda"):
        m = transvae.TransVAE(variant="large", compression_ratio=16, latent_dim=32).eval()
    x = torch.rand(B, 3, 256, 256, generator=torch.Generator().manual_seed(1)).cuda()

def run():
    te, td = {}, {}
    with torch.no_grad():
        h = m.encoder.forward_features(x, trace=te)
        mu, lv = m.encode(x)
        rec = m.decoder(mu, trace=td)
    out = dict(te)
    out["mu"] = mu
    out.update(td)
    out["recon"] = rec
    return {k: v.float().clone() for k, v in out.items()}

a, b = run(), run()
first = None
for k in a:
    d = (a[k] - b[k])
    l2 = float(d.norm() / a[k].norm().clamp_min(1e-20))
    mx = float(d.abs().max() / a[k].abs().max().clamp_min(1e-20))
    nz = int((d != 0).sum())
    if nz and first is None:
        first = k
    print(f"{k:32s} l2 {l2:9.2e}  max {mx:9.2e}  differing {nz}/{d.numel()}   |x|max {float(a[k].abs().max()):9.3e} mean {float(a[k].mean()):9.3e} std {float(a[k].std()):9.3e}")
print("first differing layer:", first)
