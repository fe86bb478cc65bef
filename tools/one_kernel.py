"""Run ONE hot kernel a few times at its BASELINE-size shape (for ncu --set full captures):
    python tools/one_kernel.py conv192 | conv192gn | gn_apply | gn_bwd | attn_fwd | attn_bwd | wgrad192 | token_norm_bwd"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepl-project_b200"))
import torch
from transvae import ops, _taps as T

which = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
if which in ("conv192", "conv192gn", "wgrad192"):
    x = rnd(B, 256, 256, 192).to(torch.bfloat16)
    w = (rnd(192, 9 * 192) * 0.02).to(torch.bfloat16)
    b = rnd(192)
    dz = rnd(B, 256, 256, 192).to(torch.bfloat16)
    # conv192gn: the same launch with the GroupNorm(32) statistics of its output taken in the epilogue (gn_sums)
    fn = (lambda: ops.mtgemm(T.plan_conv3x3(192), x, w, out_shape=(B, 256, 256, 192), bias=b, residual=dz,
                             gn_groups=32 if which == "conv192gn" else 0)) if which != "wgrad192" \
        else (lambda: ops.mtgemm_wgrad(T.plan_conv3x3(192), x, dz, 192, bias=True))
elif which in ("gn_apply", "gn_bwd"):
    x = rnd(B, 256, 256, 192).to(torch.bfloat16)
    dh = rnd(B, 256, 256, 192).to(torch.bfloat16)
    ga, be = rnd(192), rnd(192)
    s = ops.groupnorm_stats(x)
    fn = (lambda: ops.groupnorm_silu(x, ga, be, sums=s)) if which == "gn_apply" else (lambda: ops.groupnorm_bwd(x, dh, s, ga, be, add=dh))
elif which in ("attn_fwd", "attn_bwd"):
    S, C = 4096, 384
    qkv = (rnd(B, S, 3 * C) * 0.5).to(torch.bfloat16)
    o, lse = ops.attn_fwd(qkv, B, S, C, need_lse=True)
    do = torch.randn_like(o)
    tab = T.rope_table(64, 64, 1.0 / (10000 ** (torch.arange(0, 32, 2, device=dev).float() / 32)))
    fn = (lambda: ops.attn_fwd(qkv, B, S, C, need_lse=True)) if which == "attn_fwd" \
        else (lambda: ops.attn_bwd(qkv, o, do, lse, tab, B, S, C, 64, 64, 0.125))
elif which == "token_norm_bwd":
    x, dy = rnd(B * 4096, 384).to(torch.bfloat16), rnd(B * 4096, 384).to(torch.bfloat16)
    w = rnd(384)
    fn = lambda: ops.token_norm_bwd(x, w, dy, dy, 1)
elif which in ("lin384_gelu", "lin384_res"):
    M = B * 4096
    x = rnd(1, 1, M, 384).to(torch.bfloat16)
    w = (rnd(1536, 384) * 0.05).to(torch.bfloat16)
    b, rs = rnd(1536), torch.rand(M, device=dev) + 0.5
    res = rnd(1, 1, M, 1536).to(torch.bfloat16)
    fn = (lambda: ops.mtgemm(T.plan_linear(384), x, w, out_shape=(1, 1, M, 1536), bias=b, act=ops.ACT_GELU, row_scale=rs)) \
        if which == "lin384_gelu" else (lambda: ops.mtgemm(T.plan_linear(384), x, w, out_shape=(1, 1, M, 1536), bias=b, residual=res))
elif which in ("dgrad192", "dgrad192gnb"):
    # input-gradient GEMM of a ResBlock convolution, plain / with the fused GroupNorm-backward reduce pass
    x = rnd(B, 256, 256, 192).to(torch.bfloat16)
    dz = rnd(B, 256, 256, 192).to(torch.bfloat16)
    wd = (rnd(192, 9 * 192) * 0.02).to(torch.bfloat16)
    ga, be = rnd(192), rnd(192)
    s = ops.groupnorm_stats(x)
    fn = (lambda: ops.mtgemm(T.plan_conv3x3_dgrad(192), dz, wd, out_shape=(B, 256, 256, 192))) if which == "dgrad192" \
        else (lambda: ops.mtgemm(T.plan_conv3x3_dgrad(192), dz, wd, out_shape=(B, 256, 256, 192), gn_bwd=(x, s, ga, be, 32, 1e-5, True)))
elif which in ("wgrad_lin384", "wgrad_lin384t", "wgrad_lin768"):
    # weight gradients of the stage-1 / stage-2 token GEMMs (long pixel axis, small weight matrix)
    M = B * (4096 if which != "wgrad_lin768" else 1024)
    n, k = {"wgrad_lin384": (1536, 384), "wgrad_lin384t": (384, 1536), "wgrad_lin768": (3072, 768)}[which]
    a = rnd(1, 1, M, k).to(torch.bfloat16)
    dz = rnd(1, 1, M, n).to(torch.bfloat16)
    fn = lambda: ops.mtgemm_wgrad(T.plan_linear(k), a, dz, n, bias=True)
elif which in ("lin384_dgelu", "lin384_dual", "qkv384"):
    M = B * 4096
    if which == "lin384_dgelu":      # dgrad of the FFN output projection: (acc + residual) * gelu'(z)
        dy = rnd(1, 1, M, 384).to(torch.bfloat16)
        w = (rnd(1536, 384) * 0.05).to(torch.bfloat16)
        z, res = rnd(1, 1, M, 1536).to(torch.bfloat16), rnd(1, 1, M, 1536).to(torch.bfloat16)
        fn = lambda: ops.mtgemm(T.plan_linear(384), dy, w, out_shape=(1, 1, M, 1536), act=ops.ACT_GELU, act_grad_z=z, residual=res)
    elif which == "lin384_dual":     # training forward of the FFN input projection: z and gelu(z)
        x = rnd(1, 1, M, 384).to(torch.bfloat16)
        w = (rnd(1536, 384) * 0.05).to(torch.bfloat16)
        b = rnd(1536)
        fn = lambda: ops.mtgemm(T.plan_linear(384), x, w, out_shape=(1, 1, M, 1536), bias=b, act=ops.ACT_GELU, dual=True)
    else:                            # QKV projection with the LayerNorm-affine + RoPE epilogue
        x = rnd(1, 1, M, 384).to(torch.bfloat16)
        w = (rnd(1152, 384) * 0.05).to(torch.bfloat16)
        b = rnd(1152)
        tab = T.rope_table(64, 64, 1.0 / (10000 ** (torch.arange(0, 32, 2, device=dev).float() / 32)))
        fn = lambda: ops.mtgemm(T.plan_linear(384), x, w, out_shape=(1, 1, M, 1152), bias=b, rope=(tab, 384, 64, 64, 0.125))
else:
    raise SystemExit(f"unknown kernel {which}")
for _ in range(5):
    fn()
torch.cuda.synchronize()
print("ok", which)
