"""CUDA-event timing of tvae_token_norm_bwd at the three stage shapes of TransVAE-large (micro-batch 32)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepl-project_b200"))
import torch
from transvae import ops
for (M, C) in [(131072, 384), (32768, 768), (8192, 1536)]:
    x = torch.randn(M, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn(M, C, device="cuda").to(torch.bfloat16)
    add = torch.randn(M, C, device="cuda").to(torch.bfloat16)
    w = torch.randn(C, device="cuda")
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for mode in (0, 1):
        ts = []
        for _ in range(8):
            big.zero_()                       # flush L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.token_norm_bwd(x, w, dy, add, mode); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts)[len(ts) // 2]
        print(f"token_norm_bwd M={M} C={C} mode={mode}: {t:.1f} us  {M * C * 8 / t / 1e6:.2f} TB/s", flush=True)
