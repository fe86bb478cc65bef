"""Generate tests/golden/metrics_ref.pt from the UNMODIFIED reference evaluation script (authoring container only).

The reference's ``calculate_psnr`` / ``calculate_ssim`` live in a script whose other imports (lpips, matplotlib, tqdm
data loaders) are unavailable offline, so the two function definitions are taken from the script's syntax tree and
executed as they are -- nothing is copied into this repository.  Also checks the oracle restatement bit-for-bit.

    python oracle/make_golden_metrics.py
"""
from __future__ import annotations

import ast
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import transvae_oracle as O  # noqa: E402

SCRIPT = "/root/reference/transvae-implementation/transvae-implementation_patched/evaluate_transvae.py"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "metrics_ref.pt")


def reference_functions():
    tree = ast.parse(open(SCRIPT).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("calculate_psnr", "calculate_ssim")]
    assert len(keep) == 2
    ns = {"torch": torch, "F": F}
    exec(compile(ast.Module(body=keep, type_ignores=[]), SCRIPT, "exec"), ns)
    return ns["calculate_psnr"], ns["calculate_ssim"]


def main():
    ref_psnr, ref_ssim = reference_functions()
    g = torch.Generator().manual_seed(11)
    cases = {}
    for name, (b, c, h, w) in {"small": (2, 3, 40, 56), "tile_edge": (1, 3, 64, 33), "tiny": (3, 1, 7, 9)}.items():
        target = torch.rand(b, c, h, w, generator=g)
        logits = torch.randn(b, c, h, w, generator=g) * 2.0 + (target - 0.5) * 6.0      # correlated with the target
        out = {}
        for mode, f in (("sigmoid", torch.sigmoid), ("clamp", lambda t: t.clamp(0, 1)), ("none", lambda t: t)):
            r = f(logits)
            psnr = torch.tensor([ref_psnr(r[i:i + 1], target[i:i + 1]) for i in range(b)])
            ssim = torch.tensor([ref_ssim(r[i:i + 1], target[i:i + 1]) for i in range(b)])
            for i in range(b):   # the oracle restatement must agree exactly
                assert O.psnr(r[i:i + 1], target[i:i + 1]) == float(psnr[i])
                assert O.ssim(r[i:i + 1], target[i:i + 1]) == float(ssim[i])
            assert torch.equal(O.ssim(r, target, size_average=False), ref_ssim(r, target, size_average=False))
            out[mode] = dict(psnr=psnr, ssim=ssim, mse=((r - target) ** 2).mean(dim=(1, 2, 3)),
                             l1=(r - target).abs().mean(dim=(1, 2, 3)))
        cases[name] = dict(logits=logits, target=target, out=out)
    torch.save(dict(cases=cases, torch_version=torch.__version__), OUT)
    print(f"{OUT}: {os.path.getsize(OUT) / 1024:.0f} KiB; oracle == reference for psnr / ssim: OK")


if __name__ == "__main__":
    main()
