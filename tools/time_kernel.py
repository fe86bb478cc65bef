"""CUDA-event timing of one tools/one_kernel.py shape (no profiler): python tools/time_kernel.py conv192gn [B]"""
import os, runpy, sys
import torch
ns = runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "one_kernel.py"), run_name="probe")
fn = ns["fn"]
for _ in range(5):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
e0.record()
for _ in range(n):
    fn()
e1.record()
torch.cuda.synchronize()
print(f"{sys.argv[1]}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per call")
