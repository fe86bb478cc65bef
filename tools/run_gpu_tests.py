"""Run every -m gpu test in its own subprocess with a timeout (a trapped kernel poisons its CUDA context,
and a hang must not take the whole gpurun call with it).  Summary -> stdout and gpurun_out/gpu_tests.log.

    python tools/run_gpu_tests.py [pytest path or -k expression ...]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
os.makedirs("gpurun_out", exist_ok=True)
args = sys.argv[1:] or ["tests"]
col = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu", *args],
                     capture_output=True, text=True)
ids = []
for l in col.stdout.splitlines():
    if "::" in l:
        base = l.strip().split("[")[0]          # one subprocess per test function (all its parametrisations)
        if base not in ids:
            ids.append(base)
print(f"collected {len(ids)} gpu test functions", flush=True)
res = []
log = open("gpurun_out/gpu_tests.log", "w")
for tid in ids:
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", tid, "--no-header", "-p",
                            "no:cacheprovider"], capture_output=True, text=True, timeout=240)
        status = "PASS" if r.returncode == 0 else "FAIL"
        out = r.stdout + r.stderr
    except subprocess.TimeoutExpired as e:
        status, out = "TIMEOUT", (e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)
    dt = time.time() - t0
    res.append((tid, status, dt))
    print(f"{status:8s} {dt:6.1f}s {tid}", flush=True)
    log.write(f"===== {status} {tid} ({dt:.1f}s)\n")
    if status != "PASS":
        log.write(out[-6000:] + "\n")
        tail = [l for l in out.splitlines() if l.strip()][-12:]
        print("    " + "\n    ".join(tail), flush=True)
    log.flush()
bad = [r for r in res if r[1] != "PASS"]
print(f"\n{len(res) - len(bad)}/{len(res)} passed", flush=True)
sys.exit(1 if bad else 0)
