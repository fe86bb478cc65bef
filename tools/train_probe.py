"""Probe: large f16d32 training step (fwd+bwd+AdamW) time and memory on one B200."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepl-project_b200"))
import torch
import transvae
from transvae import ops
from transvae.trainer import Trainer

variant = sys.argv[1] if len(sys.argv) > 1 else "large"
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda")
torch.manual_seed(0)
with torch.device(dev):
    model = transvae.TransVAE(variant=variant, compression_ratio=16, latent_dim=32)
loss_fn = transvae.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)
tr = Trainer(model, loss_fn, lr=1e-4, accumulation_steps=1)
x = torch.rand(mb, 3, 256, 256, device=dev)
torch.cuda.reset_peak_memory_stats()
for i in range(2):
    out = tr.train_step(x)
torch.cuda.synchronize()
print("loss", {k: float(v) for k, v in out.items()}, "peak GB", torch.cuda.max_memory_allocated() / 2**30, flush=True)
if os.environ.get("TVAE_PROFILE_RANGE"):
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tr.train_step(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ops.PROFILE = []
n0 = ops.LAUNCHES
e0.record()
steps = 3
for i in range(steps):
    out = tr.train_step(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
prof, ops.PROFILE = ops.PROFILE, None
print(f"{variant} micro-batch {mb}: {ms:.1f} ms/step -> {mb / ms * 1e3:.1f} img/s fwd+bwd+opt; launches/step {(ops.LAUNCHES - n0) // steps}; "
      f"model TFLOP/s {mb / ms * 1e3 * 6187.8 / 1e3:.0f}", flush=True)
tab = {}
for name, fl, a, b, *_ in prof:
    key = name.split(" M=")[0]
    d = tab.setdefault(key, [0.0, 0.0, 0])
    d[0] += a.elapsed_time(b); d[1] += fl; d[2] += 1
tot = sum(v[0] for v in tab.values())
print(f"timed tensor kernels: {tot / steps:.1f} ms/step of {ms:.1f}")
for k, (t, fl, n) in sorted(tab.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {t / steps:9.2f} ms/step {n // steps:5d} launches  {fl / t / 1e9 if t else 0:8.1f} TFLOP/s")
