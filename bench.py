"""bench.py -- headline benchmark of the B200-native TransVAE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--res R]

N=1 workload = BASELINE.json configs[1]: TransVAE-large f16d32 encode+decode inference, bf16 compute, batch 64 at
256x256, synthetic images, random-init weights.  One "step" = mu, logvar = model.encode(x); recon = model.decode(mu)
on one batch.  N>1 (launched by torchrun, one rank per GPU): the batch dimension is sharded, every rank runs the
same per-GPU batch (weak scaling), no data-path collective; the timed region is bracketed by a barrier +
torch.cuda.synchronize(), timed on the device with CUDA events, MAX over ranks.

Printed JSON keys (one line, rank 0): metric/value/unit/..., `e2e` (same metric through the public host-to-host call
transvae.streaming.StreamedReconstructor.reconstruct: every step uploads its pinned input and downloads its
reconstruction inside the timed region, on side streams), `roofline` (tensor-pipe roofline of the dominant kernel,
tvae::mtgemm2_kernel / mtgemm_kernel, from per-launch CUDA events; `traffic` from the committed ncu capture),
`cpu_baseline` (the oracle port of the reference's PyTorch path on this box's host cores, bounded sample), `clocks`,
`gpu_launches`, `train` (BASELINE configs[2]: fwd + L1/KL loss + bwd + all-reduce + clip + AdamW on global batch 256).

`--impl reference` times the oracle port (oracle/transvae_oracle.py, a bit-exact restatement of the reference's
torch path -- the reference is pure Python and cannot travel to the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "deepl-project_b200"))

import torch  # noqa: E402

METRIC = "images_per_sec_encode_decode_256"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="large")
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="per-shape kernel table on stderr")
    ap.add_argument("--no-train", action="store_true", help="skip the fwd+bwd (BASELINE configs[2]) measurement")
    ap.add_argument("--train-global-batch", type=int, default=256)
    ap.add_argument("--train-micro-batch", type=int, default=32)
    ap.add_argument("--train-steps", type=int, default=2)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=float(p["bf16_tflops_sustained"]), hbm=float(p["hbm_gbs"]), src="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_table(prof_hbm, steps, pk):
    """Per-kernel HBM roofline of the bandwidth-bound launches: algorithmic bytes / CUDA-event time vs the measured
    copy bandwidth (MEASURED_PEAKS.json hbm_gbs)."""
    agg = {}
    for name, nbytes, a, b in prof_hbm or []:
        d = agg.setdefault(name, [0.0, 0.0, 0])
        d[0] += nbytes
        d[1] += a.elapsed_time(b)
        d[2] += 1
    out = []
    for name, (nbytes, ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        out.append({"kernel": name, "launches_per_step": n / steps, "ms_per_step": ms / steps, "achieved_gbs": gbs,
                    "frac_of_hbm_peak": gbs / pk["hbm"]})
    return out


def dist_setup(n):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif n > 1:
        raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    return rank, world, local


# ----------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_oracle_rate(variant: str, res: int, steps: int, warmup: int, batch: int):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import transvae_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.variant_config(variant)
    sd = O.init_state_dict(cfg, seed=0, mode="reference")
    x = torch.rand(batch, 3, res, res, generator=torch.Generator().manual_seed(1234))

    def step():
        with torch.no_grad():
            mu, _ = O.encode(sd, cfg, x)
            return O.decode(sd, cfg, mu)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    b = 1
    rate, ms, cores = cpu_oracle_rate(args.variant, args.res, args.steps, args.warmup, b)
    sample = f"oracle port (fp32 torch CPU), {args.variant} f16d32 encode+decode, batch {b} at {args.res}^2 per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"TransVAE-{args.variant} f16d32 encode+decode inference at {args.res}^2 (CPU, batch {b})"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local):
    import torch.distributed as dist
    import transvae
    from transvae import _lib, ops
    _lib.require_device()
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    with torch.device(dev):
        model = transvae.TransVAE(variant=args.variant, compression_ratio=16, latent_dim=32, input_resolution=args.res)
    model.eval()
    B, R = args.batch, args.res
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    x_host = torch.rand(B, 3, R, R, generator=g).pin_memory()
    x_dev = x_host.to(dev)
    out_host = torch.empty(B, 3, R, R, dtype=torch.float32).pin_memory()

    def step_resident():
        with torch.no_grad():
            mu, _ = model.encode(x_dev)
            return model.decode(mu)

    # e2e: the public host-to-host call (transvae.streaming.StreamedReconstructor): every step uploads its pinned input
    # and downloads its reconstruction; the copies run on side streams (copy engines) and overlap the kernels of the
    # neighbouring steps; the timed region ends after the last download (pipe.join before the closing event)
    from transvae.streaming import StreamedReconstructor
    pipe = StreamedReconstructor(model, dev)

    def step_e2e():
        pipe.reconstruct(x_host, out_host, next_x_host=x_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if profile:
            ops.PROFILE, ops.PROFILE_HBM = [], []
        n0 = ops.LAUNCHES
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        prof, ops.PROFILE = ops.PROFILE, None
        hbm, ops.PROFILE_HBM = ops.PROFILE_HBM, None
        if profile:
            prof = (prof, hbm)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), ops.LAUNCHES - n0, prof

    for _ in range(max(args.warmup, 3)):
        step_resident()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, launches, _ = timed(step_resident, args.steps)
    clk = clocks.stop() if rank == 0 else None
    value = B * world * args.steps / (ms / 1e3)

    # roofline of the dominant kernel from per-launch CUDA events (separate instrumented steps so the event records do
    # not perturb `value`; same stream, same inputs, directly after the timed region)
    _, _, (prof, prof_hbm) = timed(step_resident, 2, profile=True)
    pk = peaks()
    by = {}
    for name, fl, a, b in prof:
        d = by.setdefault("attn_fwd" if name.startswith("attn_fwd") else "mtgemm", [0.0, 0.0, 0])
        d[0] += fl
        d[1] += a.elapsed_time(b)
        d[2] += 1
    if args.breakdown and rank == 0:
        tab = {}
        for name, fl, a, b in prof:
            d = tab.setdefault((name, round(fl / 1e9, 1)), [0.0, 0])
            d[0] += a.elapsed_time(b)
            d[1] += 1
        print(f"{'kernel':78s} GFLOP/launch launches/step   ms/step   TFLOP/s", file=sys.stderr)
        for (name, gf), (t, n) in sorted(tab.items(), key=lambda kv: -kv[1][0]):
            print(f"{name:78s} {gf:10.1f} {n // 2:8d} {t / 2:12.3f} {gf * n / t if t else 0:9.1f}", file=sys.stderr)
    gm = by.get("mtgemm", [0.0, 1e-9, 0])
    at = by.get("attn_fwd", [0.0, 1e-9, 0])
    step_ms_prof = sum(a.elapsed_time(b) for _, _, a, b in prof) / 2
    ach = gm[0] / (gm[1] * 1e-3) / 1e12
    # DRAM traffic of the dominant launch shape from the committed ncu --set full capture (bytes per launch of that shape)
    traffic, traffic_ref = None, None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic_dominant_kernel.json")
    if os.path.exists(tpath):
        traffic_ref = json.load(open(tpath))
        traffic = traffic_ref["dram_bytes"]
    roofline = {"bound": "tensor", "kernel": "tvae::mtgemm2_kernel / mtgemm_kernel (all conv / linear launches)", "achieved": ach,
                "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"], "traffic": traffic,
                "traffic_capture": traffic_ref,
                "peak_source": pk["src"], "launches_per_step": gm[2] // 2, "ms_per_step_in_kernel": gm[1] / 2,
                "share_of_step": gm[1] / 2 / (ms / args.steps),
                "attention": {"achieved": at[0] / (at[1] * 1e-3) / 1e12, "unit": "TFLOP/s",
                              "frac": at[0] / (at[1] * 1e-3) / 1e12 / pk["tflops"], "ms_per_step_in_kernel": at[1] / 2,
                              "launches_per_step": at[2] // 2},
                "sum_of_timed_kernels_ms": step_ms_prof,
                "hbm_kernels": hbm_table(prof_hbm, 2, pk)}

    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step_e2e()
        ms_e, _, _ = timed(step_e2e, args.steps, finish=pipe.join)
        pipe.synchronize()
        e2e = {"value": B * world * args.steps / (ms_e / 1e3), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
               "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e / args.steps,
               "api": "transvae.streaming.StreamedReconstructor.reconstruct (pinned host in / out, copies on side streams)"}

    train = None
    if not args.no_train:
        train = run_train(args, model, rank, world, dev, dist, pk)

    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import transvae_oracle as O
    gflop = O.forward_flops_per_image(O.variant_config(args.variant), R) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, ms_cpu, cores = cpu_oracle_rate(args.variant, R, 2, 1, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"oracle port (fp32 torch CPU) of the same encode+decode, batch 1 at {R}^2, 1 warm-up + 2 timed steps"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"TransVAE-{args.variant} f16d32 encode+decode inference, batch {B}/GPU at {R}^2 (BASELINE configs[1])",
                   "per_gpu_batch": B, "resolution": R, "parallelism": f"batch-sharded x{world}, no collective",
                   "l2_policy": "activations per step (>2 GB) exceed the 126 MB L2; no flush needed",
                   "gflop_per_image": gflop},
        "model_tflops": value * gflop / 1e3, "model_frac_of_peak": value * gflop / 1e3 / (pk["tflops"] * world),
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "train": train,
    }
    print(json.dumps(line), flush=True)


def run_train(args, model, rank, world, dev, dist, pk):
    """BASELINE configs[2]: stage-1 training step (fwd + L1/KL loss + bwd + gradient all-reduce + clip + AdamW) on a fixed
    global batch (strong scaling: per-GPU batch = global / N, processed in micro-batches).  One step = one optimiser
    step over the whole global batch; timed on the device, MAX over ranks."""
    import transvae
    from transvae import ops
    from transvae.trainer import Trainer
    per_gpu = max(1, args.train_global_batch // world)
    mb = min(args.train_micro_batch, per_gpu)
    accum = max(1, per_gpu // mb)
    loss_fn = transvae.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)
    tr = Trainer(model, loss_fn, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.0, grad_clip=1.0, accumulation_steps=accum)
    g = torch.Generator(device="cpu").manual_seed(4321 + rank)
    xs = [torch.rand(mb, 3, args.res, args.res, generator=g).to(dev) for _ in range(min(accum, 2))]

    def step():
        for i in range(accum):
            out = tr.train_step(xs[i % len(xs)])
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.cuda.reset_peak_memory_stats()
    step()                                  # warm-up (allocator, NCCL channels)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.LAUNCHES
    e0.record()
    for _ in range(args.train_steps):
        out = step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / args.train_steps
    # one instrumented micro-step (forward + backward, no optimiser step) for the per-kernel tables
    ops.PROFILE, ops.PROFILE_HBM = [], []
    tr._micro = 0
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    tr.train_step(xs[0])
    e3.record()
    barrier()
    prof_t, ops.PROFILE = ops.PROFILE, None
    prof_h, ops.PROFILE_HBM = ops.PROFILE_HBM, None
    tens = {}
    for name, fl, a, b in prof_t:
        key = "attn_fwd" if name.startswith("attn_fwd") else "attn_bwd" if name.startswith("attn_bwd") else \
            "wgrad" if name.startswith("wgrad") else "mtgemm (fwd + dgrad)"
        d = tens.setdefault(key, [0.0, 0.0, 0])
        d[0] += fl
        d[1] += a.elapsed_time(b)
        d[2] += 1
    tensor_tab = [{"kernel": k, "launches": n, "ms": t_ms, "achieved_tflops": fl / (t_ms * 1e-3) / 1e12,
                   "frac_of_tensor_peak": fl / (t_ms * 1e-3) / 1e12 / pk["tflops"]}
                  for k, (fl, t_ms, n) in sorted(tens.items(), key=lambda kv: -kv[1][1])]
    if args.breakdown and rank == 0:
        tab = {}
        for name, fl, a, b in prof_t:
            d = tab.setdefault((name, round(fl / 1e9, 1)), [0.0, 0])
            d[0] += a.elapsed_time(b)
            d[1] += 1
        print(f"\n[training micro-step, micro-batch {mb}] {'kernel':60s} GFLOP/launch launches   ms   TFLOP/s", file=sys.stderr)
        for (name, gf), (tt, n) in sorted(tab.items(), key=lambda kv: -kv[1][0]):
            print(f"{name:78s} {gf:10.1f} {n:8d} {tt:12.3f} {gf * n / tt if tt else 0:9.1f}", file=sys.stderr)
    micro = {"micro_batch": mb, "ms": e2.elapsed_time(e3), "tensor_kernels": tensor_tab, "hbm_kernels": hbm_table(prof_h, 1, pk)}
    imgs = mb * accum * world
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import transvae_oracle as O
    gflop = 3.0 * O.forward_flops_per_image(O.variant_config(args.variant), args.res) / 1e9
    rate = imgs / ms * 1e3
    return {"metric": "images_per_sec_fwd_bwd_256", "value": rate, "unit": UNIT, "ms_per_step": ms, "scaling": "strong",
            "global_batch": imgs, "per_gpu_micro_batch": mb, "accumulation": accum, "steps": args.train_steps,
            "loss": float(out["total"]), "gflop_per_image_fwd_bwd": gflop, "model_tflops": rate * gflop / 1e3,
            "model_frac_of_peak": rate * gflop / 1e3 / (pk["tflops"] * world),
            "gpu_launches_per_step": (ops.LAUNCHES - n0) // args.train_steps,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
            "includes": "fwd, L1+KL loss, bwd, bucketed NCCL all-reduce overlapped with bwd, clip, fused AdamW",
            "profiled_micro_step": micro}


def main():
    args = parse()
    if args.impl == "reference":
        # CPU arm: rank 0 alone runs and prints; the other ranks exit 0 without work (no process group needed)
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, world, local = dist_setup(args.gpus)
    try:
        run_ours(args, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
