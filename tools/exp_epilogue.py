"""Experiment: where does the mtgemm epilogue time go?  (GPU box only)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepl-project_b200"))
import torch
from transvae import ops, _taps as T

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

dev = "cuda"
for (M, N, K) in [(262144, 1536, 384), (65536, 3072, 768), (16384, 1536, 6144), (262144, 384, 384)]:
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).to(torch.bfloat16)
    plan = T.plan_linear(K)
    out = torch.empty(1, 1, M, N, dtype=torch.bfloat16, device=dev)
    gf = 2.0 * M * N * K / 1e9
    x4 = x.reshape(1, 1, M, K)
    t = {}
    t["tma_store"] = timeit(lambda: ops.mtgemm(plan, x4, w, out=out))
    t["tma_store+bias"] = timeit(lambda: ops.mtgemm(plan, x4, w, out=out, bias=b))
    t["tma_store+bias+gelu"] = timeit(lambda: ops.mtgemm(plan, x4, w, out=out, bias=b, act=ops.ACT_GELU))
    t["tma_store+res"] = timeit(lambda: ops.mtgemm(plan, x4, w, out=out, residual=res.reshape(1, 1, M, N)))
    dummy = torch.empty(8, device=dev)
    t["no_store(out_n=0)"] = timeit(lambda: ops.mtgemm(plan, x4, w, out_f32=dummy, out_n=0))
    print(f"M={M} N={N} K={K} ({gf:.0f} GF): " + "  ".join(f"{k}={v:.3f}ms ({gf/v:.0f} TF/s)" for k, v in t.items()), flush=True)
