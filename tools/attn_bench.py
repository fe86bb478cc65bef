"""Attention forward / backward micro-benchmark (GPU box only): the three TransVAE-large shapes at batch 64."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepl-project_b200"))
import torch
from transvae import _lib
if os.environ.get("TVAE_LIB"):
    _lib.LIB_PATH = os.environ["TVAE_LIB"]      # A/B builds of the library
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
    if "--bwd" in sys.argv:
        from transvae import _taps as T
        H = W = int(math.isqrt(S))
        inv = 1.0 / (10000 ** (torch.arange(0, 32, 2, device="cuda").float() / 32))
        tab = T.rope_table(H, W, inv)
        o, lse = ops.attn_fwd(qkv, B, S, C, need_lse=True)
        do = torch.randn_like(o)
        tb = timeit(lambda: ops.attn_bwd(qkv, o, do, lse, tab, B, S, C, H, W, 0.125))
        print(f"attn_bwd B={B} S={S} C={C}: {tb:.3f} ms  {2.5 * fl / tb / 1e9:.0f} TFLOP/s (incl. delta + rope_bwd kernels)", flush=True)
