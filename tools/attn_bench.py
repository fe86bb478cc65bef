"""Attention forward / backward micro-benchmark (GPU box only): the three TransVAE-large shapes at batch 64."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepl-project_b200"))
import torch
from transvae import ops

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for (S, C) in [(4096, 384), (1024, 768), (256, 1536)]:
    qkv = (torch.randn(B, S, 3 * C, device="cuda") * 0.5).to(torch.bfloat16)
    fl = 4.0 * B * S * S * C
    t = timeit(lambda: ops.attn_fwd(qkv, B, S, C, need_lse=True))
    print(f"attn_fwd B={B} S={S} C={C}: {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s", flush=True)
