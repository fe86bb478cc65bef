"""Phase trace of the attention forward kernel (GPU box only).  Build the trace library first, here or on the box:
    python tools/att_trace.py build
then on the GPU:  python tools/att_trace.py run
Prints, per K/V block, the SM-clock time a softmax warp spent in each phase (CTA 0, both Q tiles)."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "deepl-project_b200")
LIB = os.path.join(PKG, "libtransvae_trace.so")
if sys.argv[1] == "build":
    srcs = ["runtime.cu", "mtgemm.cu", "mtgemm2.cu", "wgrad.cu", "elementwise.cu", "attention.cu", "attention_bwd.cu", "backward.cu", "metrics.cu", "api.cu"]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "--use_fast_math", "-DTVAE_ATT_TRACE", "-shared", "-o", LIB, "-cudart", "shared"] + [os.path.join(PKG, "csrc", s) for s in srcs]
    subprocess.check_call(cmd)
    print(LIB)
    sys.exit(0)
sys.path.insert(0, PKG)
import torch
from transvae import _lib
_lib.LIB_PATH = LIB
from transvae import ops
B, S, C = 64, 4096, 384
qkv = (torch.randn(B, S, 3 * C, device="cuda") * 0.5).to(torch.bfloat16)
for _ in range(3):
    ops.attn_fwd(qkv, B, S, C, need_lse=True)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (2 * 64 * 8))()
lib = _lib.load()
lib.tvae_debug_att_trace.restype = ctypes.c_int
assert lib.tvae_debug_att_trace(buf) == 0
names = ["wait_s", "ld+free", "max", "wait_o", "exp", "(loop)"]
for t in range(2):
    print(f"tile {t}: block  start  " + " ".join(f"{n:>8s}" for n in names) + "   total")
    base = buf[(t * 64) * 8]
    prev_end = None
    for j in range(int(os.environ.get("TRACE_ROWS", "32"))):
        st = [buf[(t * 64 + j) * 8 + k] for k in range(6)]
        d = [st[k + 1] - st[k] for k in range(5)]
        nxt = buf[(t * 64 + j + 1) * 8] if j < 31 else st[5]
        print(f"        {j:5d} {st[0] - base:6d}  " + " ".join(f"{x:8d}" for x in d) + f" {nxt - st[5]:8d}   {nxt - st[0]:6d}")
