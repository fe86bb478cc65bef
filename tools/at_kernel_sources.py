"""Attribute the torch-native (at::) kernels of one training micro-step to the Python lines that launch them
(torch.profiler with stacks): python tools/at_kernel_sources.py [variant] [micro_batch]"""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepl-project_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import transvae
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
for _ in range(2):
    tr.train_step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True,
             experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
    tr.train_step(x)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    name = ev.name
    if not name.startswith("aten::"):
        continue
    # only leaf aten ops (those that own kernels): self device time
    t = ev.self_device_time_total
    if t <= 0:
        continue
    stack = ev.stack or []
    where = next((s for s in stack if "deepl-project_b200" in s or "transvae" in s), None)
    if where is None:
        where = "(autograd engine) " + str(ev.input_shapes)[:90]
    where = where.replace(ROOT + "/", "")
    k = (name, where)
    agg[k][0] += 1
    agg[k][1] += t / 1e3
tot = sum(v[1] for v in agg.values())
print(f"aten ops with device time: {sum(v[0] for v in agg.values())} calls, {tot:.2f} ms")
for (name, where), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{ms:8.3f} ms {n:5d} x {name:28s} {where}")
