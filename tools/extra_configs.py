"""BASELINE configs[3] / [4]: resolution extrapolation (512^2, 1024^2) and the giant variant -- run-and-report on one B200.
Prints one JSON line per configuration (inference img/s; giant also one training micro-step)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepl-project_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import transvae, transvae_oracle as O
from transvae import ops
from transvae.trainer import Trainer

dev = torch.device("cuda")

def infer(variant, res, B, steps=3):
    torch.manual_seed(0)
    with torch.device(dev):
        m = transvae.TransVAE(variant=variant, compression_ratio=16, latent_dim=32).eval()
    x = torch.rand(B, 3, res, res, device=dev)
    with torch.no_grad():
        for _ in range(2):
            mu, _ = m.encode(x); r = m.decode(mu)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            mu, _ = m.encode(x); r = m.decode(mu)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    gf = O.forward_flops_per_image(O.variant_config(variant), res) / 1e9
    print(json.dumps({"config": f"{variant} f16d32 encode+decode @{res}^2 batch {B}", "images_per_s": B / ms * 1e3, "ms_per_step": ms,
                      "gflop_per_image": gf, "model_tflops": B / ms * gf, "finite": bool(torch.isfinite(r).all()),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
    return m

which = sys.argv[1:] or ["512", "1024", "giant"]
if "512" in which:
    infer("large", 512, 16); torch.cuda.empty_cache()
if "1024" in which:
    infer("large", 1024, 4); torch.cuda.empty_cache()
if "giant" in which:
    m = infer("giant", 256, 16)
    loss_fn = transvae.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)
    tr = Trainer(m, loss_fn, lr=1e-4)
    mb = 8
    x = torch.rand(mb, 3, 256, 256, device=dev)
    tr.train_step(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = tr.train_step(x); out = tr.train_step(x); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    gf = 3 * O.forward_flops_per_image(O.variant_config("giant"), 256) / 1e9
    print(json.dumps({"config": f"giant f16d32 training micro-step batch {mb} @256^2 (fwd+bwd+AdamW)", "images_per_s": mb / ms * 1e3,
                      "ms_per_step": ms, "model_tflops": mb / ms * gf, "loss": float(out["total"]),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
