"""bench.py -- headline benchmark of the B200-native TransVAE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME] [--breakdown-json FILE]

Default workload (`--config train256`) = BASELINE.json configs[2]: TransVAE-large f16d32 stage-1 training step
(forward, L1 + 1e-8 KL loss, backward, gradient all-reduce, clip, AdamW) on a GLOBAL batch of 256 images at 256x256,
bf16 compute / fp32 master weights, synthetic images, random-init weights.  One "step" = one optimiser step over the
whole global batch: every rank processes 256 / N images in micro-batches of 32 (gradient accumulation, the all-reduce
fires on the last micro-step only and overlaps its backward) -- strong scaling, the only curve of this path that
contains a collective.  Launched by torchrun for N > 1 (one rank per GPU, NCCL); the timed region is bracketed by a
barrier + torch.cuda.synchronize(), timed on the device with CUDA events, MAX over ranks.

Other workloads: `infer256` (BASELINE configs[1]: encode+decode, batch 64 / GPU), `extrap512` / `extrap1024`
(configs[3]: batch-sharded inference at 512^2 / 1024^2, batch 8 / 2 per GPU), `giant` (configs[4]: the 4.8 B-parameter
variant, DDP training, micro-batch 8).  Inference workloads shard the batch with no collective (weak scaling).

Printed JSON (ONE line, rank 0, < 1.5 kB): metric / value / unit / ..., `e2e` (the same metric through the public
host-to-host call -- Trainer.train_step_host / StreamedReconstructor.reconstruct -- with the pinned-host upload of every
micro-batch and the download of the result inside the timed region), `roofline` (tensor-pipe roofline of the dominant
kernel, tvae::mtgemm2_kernel / mtgemm_kernel, from per-launch CUDA events of instrumented micro-steps; `frac` counts the
reference's algorithmic FLOPs, `frac_executed` only the MACs the kernel really executes -- the nearest-2x upsample
convolution runs in its 4-phase 2x2 form; `traffic` from the committed ncu capture), `cpu_baseline`, `clocks`,
`gpu_launches`, and at N = 1 the secondary legs `infer` (configs[1]) and `gpu_stock_torch` (the UNMODIFIED reference on
the same B200 under bf16 autocast: cuDNN / cuBLAS / SDPA -- the practical comparator).  Per-kernel tables go to
`--breakdown-json` (and to stderr with --breakdown), never into the line.

`--impl reference` times the reference's own CPU implementation of the same workload on the host cores: the unmodified
package from baseline/_ref (kind "reference"; pip-installed there, see baseline/reference_runner.py) or, if that is
absent, the oracle port (kind "port"); each step is a bounded sample (batch 1) of the workload.
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

UNIT = "images/s"
CONFIGS = {
    "train256": dict(kind="train", variant="large", res=256, global_batch=256, micro=32, baseline_cfg=2),
    "infer256": dict(kind="infer", variant="large", res=256, batch=64, baseline_cfg=1),
    "extrap512": dict(kind="infer", variant="large", res=512, batch=8, baseline_cfg=3),
    "extrap1024": dict(kind="infer", variant="large", res=1024, batch=2, baseline_cfg=3),
    "giant": dict(kind="train", variant="giant", res=256, global_batch=256, micro=8, baseline_cfg=4),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="train256", choices=sorted(CONFIGS))
    ap.add_argument("--global-batch", type=int, default=None, help="training: images per optimiser step over all ranks")
    ap.add_argument("--micro-batch", type=int, default=None, help="training: images per forward/backward per GPU")
    ap.add_argument("--batch", type=int, default=None, help="inference: per-GPU batch")
    ap.add_argument("--grad-comm", default="fp32", choices=["fp32", "bf16"], help="dtype of the gradient all-reduce")
    ap.add_argument("--bucket-mb", type=int, default=64)
    ap.add_argument("--checkpointing", action="store_true", help="training: per-block activation recompute")
    ap.add_argument("--no-legs", action="store_true", help="skip the N = 1 secondary legs (infer / stock torch / CPU)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="per-shape kernel tables on stderr")
    ap.add_argument("--breakdown-json", default=None, help="write the per-kernel tables (and the full record) here")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=float(p["bf16_tflops_sustained"]), hbm=float(p["hbm_gbs"]), src="MEASURED_PEAKS.json (sustained)")
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


def hbm_table(prof_hbm, per, pk):
    """Per-kernel HBM roofline of the bandwidth-bound launches: algorithmic bytes / CUDA-event time vs the measured copy
    bandwidth (MEASURED_PEAKS.json hbm_gbs).  `per` = number of profiled units (steps / micro-steps)."""
    agg = {}
    for name, nbytes, a, b in prof_hbm or []:
        d = agg.setdefault(name, [0.0, 0.0, 0])
        d[0] += nbytes
        d[1] += a.elapsed_time(b)
        d[2] += 1
    out = []
    for name, (nbytes, ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        out.append({"kernel": name, "launches": n / per, "ms": round(ms / per, 4), "achieved_gbs": round(gbs, 1),
                    "frac_of_hbm_peak": round(gbs / pk["hbm"], 4)})
    return out


def tensor_tables(prof, per, pk):
    """(per-class table, per-shape table) of the tensor-core launches from (tag, algorithmic flops, e0, e1) records."""
    cls, shapes = {}, {}
    for name, fl, a, b, _ex in prof or []:
        key = "attn_fwd" if name.startswith("attn_fwd") else "attn_bwd" if name.startswith("attn_bwd") else \
            "wgrad" if name.startswith("wgrad") else \
            "mtgemm dgrad + GroupNorm-backward reduce pass (fused epilogue)" if name.endswith("+gn_bwd_reduce") else \
            "mtgemm (fwd + dgrad)"
        t = a.elapsed_time(b)
        d = cls.setdefault(key, [0.0, 0.0, 0])
        d[0] += fl
        d[1] += t
        d[2] += 1
        s = shapes.setdefault((name, round(fl / 1e9, 1)), [0.0, 0])
        s[0] += t
        s[1] += 1
    ctab = [{"kernel": k, "launches": n / per, "ms": round(t / per, 3), "achieved_tflops": round(fl / (t * 1e-3) / 1e12, 1),
             "frac_of_tensor_peak": round(fl / (t * 1e-3) / 1e12 / pk["tflops"], 4)}
            for k, (fl, t, n) in sorted(cls.items(), key=lambda kv: -kv[1][1]) if t > 0]
    stab = [{"launch": k[0], "gflop": k[1], "launches": n / per, "ms": round(t / per, 4),
             "tflops": round(k[1] * n / t, 1) if t else 0.0}
            for k, (t, n) in sorted(shapes.items(), key=lambda kv: -kv[1][0])]
    return ctab, stab


def print_shapes(title, stab):
    print(f"\n[{title}] {'launch':78s} GFLOP/launch  launches      ms   TFLOP/s", file=sys.stderr)
    for r in stab:
        print(f"{r['launch']:88s} {r['gflop']:10.1f} {r['launches']:8.1f} {r['ms']:10.3f} {r['tflops']:9.1f}", file=sys.stderr)


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


def flops_per_image(variant: str, res: int) -> float:
    """Algorithmic forward GFLOP / image as the reference's modules count them (oracle FLOP model)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import transvae_oracle as O
    return O.forward_flops_per_image(O.variant_config(variant), res) / 1e9


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own torch code on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg: dict, steps: int, warmup: int):
    """(img/s, ms/step, cores, kind, sample) of the reference's CPU implementation on a batch-1 sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_runner as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    train = cfg["kind"] == "train"
    res, variant = cfg["res"], cfg["variant"]
    x = torch.rand(1, 3, res, res, generator=torch.Generator().manual_seed(1234))
    model, pkg = R.build(variant, patched=train, seed=0)
    if model is not None:
        kind = "reference"
        if train:
            model.train()
            loss_fn = pkg.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)

            def step():
                model.zero_grad(set_to_none=True)
                rec, mu, lv = model(x)
                loss_fn(rec, x, mu, lv)["total"].backward()
        else:
            model.eval()

            def step():
                with torch.no_grad():
                    mu, _ = model.encode(x)
                    model.decode(mu)
    else:
        import transvae_oracle as O
        kind = "port"
        ocfg = O.variant_config(variant)
        sd = O.init_state_dict(ocfg, seed=0, mode="reference")
        if train:
            sdg = {k: v.clone().requires_grad_("inv_freq" not in k) for k, v in sd.items()}
            eps = torch.randn(1, 32, res // 16, res // 16, generator=torch.Generator().manual_seed(5))

            def step():
                for v in sdg.values():
                    v.grad = None
                rec, mu, lv, _ = O.forward(sdg, ocfg, x, eps, patched=True)
                O.loss_l1_kl(rec, x, mu, lv, patched=True)["total"].backward()
        else:
            def step():
                with torch.no_grad():
                    mu, _ = O.encode(sd, ocfg, x)
                    O.decode(sd, ocfg, mu)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    what = "fwd+loss+bwd (patched)" if train else "encode+decode"
    sample = (f"{'baseline/_ref' if kind == 'reference' else 'oracle port'}, fp32 CPU, {what}, batch 1 @{res}^2 per step, "
              f"{warmup}+{steps} steps")
    return steps / dt, dt / steps * 1e3, torch.get_num_threads(), kind, sample


def workload_string(name: str, cfg: dict) -> str:
    if cfg["kind"] == "train":
        return (f"TransVAE-{cfg['variant']} f16d32 train fwd+bwd L1+KL, batch {cfg['global_batch']} @{cfg['res']}^2, DDP "
                f"(configs[{cfg['baseline_cfg']}])")
    return (f"TransVAE-{cfg['variant']} f16d32 encode+decode @{cfg['res']}^2, batch {cfg['batch']}/GPU "
            f"(configs[{cfg['baseline_cfg']}])")


def metric_name(cfg: dict) -> str:
    return f"images_per_sec_fwd_bwd_{cfg['res']}" if cfg["kind"] == "train" else f"images_per_sec_encode_decode_{cfg['res']}"


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    rate, ms, cores, kind, sample = cpu_reference_rate(cfg, args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": metric_name(cfg), "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if cfg["kind"] == "train" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.config, cfg)},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# stock torch on the same B200: the unmodified reference modules under bf16 autocast (cuDNN / cuBLAS / SDPA)
# ----------------------------------------------------------------------------------------------------------------------
def stock_torch_leg(cfg: dict, dev) -> dict:
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import reference_runner as R
    out = {}        # the unmodified reference package on this GPU: torch bf16 autocast (cuDNN / cuBLAS / SDPA), fused AdamW
    if not R.available(True):
        return {"unavailable": "baseline/_ref not installed"}
    torch.backends.cudnn.benchmark = True
    res, variant = cfg["res"], cfg["variant"]

    def timed(fn, warm, n):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    try:
        with torch.device(dev):
            model, pkg = R.build(variant, patched=True, seed=0)
        model = model.to(memory_format=torch.channels_last)          # train_working.py:478
        B = 32
        x = torch.rand(B, 3, res, res, device=dev).contiguous(memory_format=torch.channels_last)
        model.eval()

        def infer():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                mu, _ = model.encode(x)
                model.decode(mu)
        ms = timed(infer, 2, 3)
        out["infer"] = {"value": round(B / ms * 1e3, 2), "ms": round(ms, 1), "batch": B}
        model.train()
        loss_fn = pkg.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0).to(dev)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.0, fused=True)
        Bt = 16
        xt = x[:Bt]

        def train():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                rec, mu, lv = model(xt)
            losses = loss_fn(rec.float(), xt, mu.float(), lv.float())       # train_working.py:353-362
            losses["total"].backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
        ms = timed(train, 2, 3)
        out["train"] = {"value": round(Bt / ms * 1e3, 2), "ms": round(ms, 1), "batch": Bt}
    except Exception as e:  # noqa: BLE001  (OOM or a missing cuDNN path must not kill the bench line)
        out["error"] = f"{type(e).__name__}: {str(e)[:160]}"
    finally:
        model = opt = None
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args, rank, world, local):
        import torch.distributed as dist
        self.args, self.rank, self.world, self.local, self.dist = args, rank, world, local, dist
        self.dev = torch.device("cuda", local)
        self.pk = peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, finish=None):
        """K calls of fn bracketed by barrier + synchronize, CUDA events, MAX over ranks -> (ms total, launches)."""
        from transvae import ops
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.LAUNCHES
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        self.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0]), ops.LAUNCHES - n0

    def profiled(self, fn, reps):
        from transvae import ops
        self.barrier()
        ops.PROFILE, ops.PROFILE_HBM = [], []
        for _ in range(reps):
            fn()
        self.barrier()
        prof, ops.PROFILE = ops.PROFILE, None
        hbm, ops.PROFILE_HBM = ops.PROFILE_HBM, None
        return prof, hbm


def traffic_capture():
    for name in ("r2_traffic_dominant_kernel.json", "r1_traffic_dominant_kernel.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            t = json.load(open(p))
            return t.get("dram_bytes"), name
    return None, None


def roofline_of(prof, per, step_ms, pk):
    """Roofline of the dominant kernel (all tvae::mtgemm* forward / input-gradient launches) from the profiled records:
    `frac` counts the reference's algorithmic FLOPs, `frac_executed` only the MACs that were really executed.  The
    ResBlock input-gradient launches whose epilogue also runs the reduce pass of the GroupNorm backward (HBM-side work of
    another operator riding on the GEMM: x tile staged, two sums per channel) and whose timed region contains the two
    tile-sum launches are a class of their own in the tables; `frac_with_fused_gn_bwd` is the same fraction with them
    included."""
    fl = ex = t = 0.0
    fl_g = t_g = 0.0
    for name, f, a, b, fe in prof:
        if name.startswith(("attn_", "wgrad")):
            continue
        dt = a.elapsed_time(b)
        if name.endswith("+gn_bwd_reduce"):
            fl_g += f
            t_g += dt
            continue
        fl += f
        ex += fe
        t += dt
    ach = fl / (t * 1e-3) / 1e12 if t > 0 else 0.0
    ach_ex = ex / (t * 1e-3) / 1e12 if t > 0 else 0.0
    ach_all = (fl + fl_g) / ((t + t_g) * 1e-3) / 1e12 if t + t_g > 0 else 0.0
    traffic, src = traffic_capture()
    out = {"bound": "tensor", "kernel": "tvae::mtgemm2_kernel (fwd+dgrad)",
           "achieved": round(ach, 1), "peak": pk["tflops"], "unit": "TFLOP/s", "frac": round(ach / pk["tflops"], 4),
           "frac_executed": round(ach_ex / pk["tflops"], 4), "traffic": traffic,
           "share_of_step": round((t + t_g) / per / step_ms, 3) if step_ms else None}
    if t_g > 0:
        out["frac_with_fused_gn_bwd"] = round(ach_all / pk["tflops"], 4)
    return out


def run_train(C: Ctx, name: str, cfg: dict) -> dict:
    import transvae
    from transvae import ops
    from transvae.trainer import Trainer
    args, dev, world, rank = C.args, C.dev, C.world, C.rank
    gb = args.global_batch or cfg["global_batch"]
    per_gpu = max(1, gb // world)
    mb = min(args.micro_batch or cfg["micro"], per_gpu)
    accum = max(1, per_gpu // mb)
    imgs = mb * accum * world
    res = cfg["res"]
    torch.manual_seed(0)
    with torch.device(dev):
        model = transvae.TransVAE(variant=cfg["variant"], compression_ratio=16, latent_dim=32, input_resolution=res)
    if args.checkpointing:
        model.enable_gradient_checkpointing()
    loss_fn = transvae.TransVAELoss(l1_weight=1.0, lpips_weight=0.0, kl_weight=1e-8, vf_weight=0.0, gan_weight=0.0)
    tr = Trainer(model, loss_fn, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.0, grad_clip=1.0, accumulation_steps=accum,
                 bucket_bytes=args.bucket_mb << 20, grad_comm=torch.bfloat16 if args.grad_comm == "bf16" else torch.float32)
    g = torch.Generator(device="cpu").manual_seed(4321 + rank)
    x_host = [torch.rand(mb, 3, res, res, generator=g).pin_memory() for _ in range(accum)]
    xs = [t.to(dev) for t in x_host]

    def step_resident():
        for i in range(accum):
            out = tr.train_step(xs[i])
        return out

    def step_host():
        for i in range(accum):
            tr.train_step_host(x_host[i], next_images_host=x_host[(i + 1) % accum])
        torch.cuda.current_stream().synchronize()        # the loss terms of the step are on the host (`loss.item()`)

    torch.cuda.reset_peak_memory_stats()
    W = max(args.warmup, 3)
    for _ in range(W):
        out = step_resident()
    clocks = ClockSampler(C.local)
    if rank == 0:
        clocks.start()
    ms, launches = C.timed(step_resident, args.steps)
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms / args.steps
    value = imgs / ms_step * 1e3
    loss = float(out["total"])

    # exposed communication / optimiser tail (device events inside the trainer), then instrumented micro-steps
    tr.timing = []
    C.timed(step_resident, 1)
    tail = tr.timing
    tr.timing = None
    comm = None
    if tail:
        comm = {"payload_gb": round(tr.buckets.numel * (2 if tr.buckets.comm_g is not None else 4) / 1e9, 2),
                "buckets": len(tr.buckets.buckets), "overlap": bool(tr.buckets.overlap), "registered": bool(tr.buckets.registered), "exposed_tail_ms": round(tail[-1][0].elapsed_time(tail[-1][1]), 2),
                "optimizer_ms": round(tail[-1][1].elapsed_time(tail[-1][2]), 2)}
    reps = min(accum, 2)
    tr._micro = 0          # profile plain micro-steps (no optimiser step inside: accumulation position 0 ..)
    prof, hbm = C.profiled(lambda: [tr.train_step(xs[i]) for i in range(reps)] if accum > reps else step_resident(), 1)
    if accum > reps:       # finish the interrupted accumulation cycle so that the trainer state stays consistent
        for i in range(reps, accum):
            tr.train_step(xs[i])
    per = reps if accum > reps else accum
    algo_gf = flops_per_image(cfg["variant"], res)
    micro_ms = sum(a.elapsed_time(b) for _, _, a, b, _e in prof) / per + sum(a.elapsed_time(b) for _, _, a, b in hbm) / per
    roof = roofline_of(prof, per, ms_step / accum, C.pk)
    ctab, stab = tensor_tables(prof, per, C.pk)
    htab = hbm_table(hbm, per, C.pk)
    roof["attn_bwd_frac"] = next((r["frac_of_tensor_peak"] for r in ctab if r["kernel"] == "attn_bwd"), None)
    roof["attn_fwd_frac"] = next((r["frac_of_tensor_peak"] for r in ctab if r["kernel"] == "attn_fwd"), None)
    roof["wgrad_frac"] = next((r["frac_of_tensor_peak"] for r in ctab if r["kernel"] == "wgrad"), None)
    if args.breakdown and rank == 0:
        print_shapes(f"training micro-step, micro-batch {mb}", stab)

    e2e = None
    if not args.no_e2e:
        step_host()
        ms_e, _ = C.timed(step_host, args.steps)
        e2e = {"value": round(imgs / (ms_e / args.steps) * 1e3, 2), "unit": UNIT,
               "h2d_bytes_per_step": accum * mb * 3 * res * res * 4, "d2h_bytes_per_step": accum * 6 * 4,
               "ms_per_step": round(ms_e / args.steps, 2), "api": "Trainer.train_step_host"}

    gf3 = 3.0 * algo_gf
    line = {
        "metric": metric_name(cfg), "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": round(ms_step, 2), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": workload_string(name, cfg), "global_batch": imgs, "micro_batch": mb, "accumulation": accum,
                   "parallelism": f"dp{world}", "grad_comm": args.grad_comm, "l2": "working set >> 126 MB L2, no flush",
                   "gflop_per_image": round(gf3, 1), **({"checkpointing": True} if args.checkpointing else {})},
        "model_frac_of_peak": round(value * gf3 / 1e3 / (C.pk["tflops"] * world), 4),
        "loss": round(loss, 4), "mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "comm": comm,
    }
    extra = {"tensor_kernels": ctab, "hbm_kernels": htab, "shapes": stab, "profiled_micro_step_kernel_ms": round(micro_ms, 2)}
    # free the training state before the secondary legs
    tr = model = xs = None
    torch.cuda.empty_cache()
    return line, extra


def run_infer(C: Ctx, name: str, cfg: dict, steps: int, warmup: int, e2e_on: bool = True):
    import transvae
    from transvae.streaming import StreamedReconstructor
    args, dev, world, rank = C.args, C.dev, C.world, C.rank
    B, R = args.batch or cfg["batch"], cfg["res"]
    torch.manual_seed(0)
    with torch.device(dev):
        model = transvae.TransVAE(variant=cfg["variant"], compression_ratio=16, latent_dim=32, input_resolution=R)
    model.eval()
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    x_host = torch.rand(B, 3, R, R, generator=g).pin_memory()
    x_dev = x_host.to(dev)
    out_host = torch.empty(B, 3, R, R, dtype=torch.float32).pin_memory()

    def step_resident():
        with torch.no_grad():
            mu, _ = model.encode(x_dev)
            return model.decode(mu)

    pipe = StreamedReconstructor(model, dev)

    def step_e2e():
        pipe.reconstruct(x_host, out_host, next_x_host=x_host)

    W = max(warmup, 3)
    for _ in range(W):
        step_resident()
    clocks = ClockSampler(C.local)
    if rank == 0:
        clocks.start()
    ms, launches = C.timed(step_resident, steps)
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms / steps
    value = B * world / ms_step * 1e3
    prof, hbm = C.profiled(step_resident, 2)
    algo_gf = flops_per_image(cfg["variant"], R)
    roof = roofline_of(prof, 2, ms_step, C.pk)
    ctab, stab = tensor_tables(prof, 2, C.pk)
    roof["attn_fwd_frac"] = next((r["frac_of_tensor_peak"] for r in ctab if r["kernel"] == "attn_fwd"), None)
    if args.breakdown and rank == 0:
        print_shapes(f"inference step, batch {B} at {R}^2", stab)
    e2e = None
    if e2e_on and not args.no_e2e:
        for _ in range(2):
            step_e2e()
        ms_e, _ = C.timed(step_e2e, steps, finish=pipe.join)
        pipe.synchronize()
        e2e = {"value": round(B * world / (ms_e / steps) * 1e3, 2), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
               "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": round(ms_e / steps, 2),
               "api": "StreamedReconstructor.reconstruct"}
    line = {
        "metric": metric_name(cfg), "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": W,
        "ms_per_step": round(ms_step, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": workload_string(name, cfg), "per_gpu_batch": B, "resolution": R,
                   "parallelism": f"batch-sharded x{world}, no collective",
                   "l2": "working set >> 126 MB L2, no flush", "gflop_per_image": round(algo_gf, 1)},
        "model_frac_of_peak": round(value * algo_gf / 1e3 / (C.pk["tflops"] * world), 4),
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roof,
    }
    extra = {"tensor_kernels": ctab, "hbm_kernels": hbm_table(hbm, 2, C.pk), "shapes": stab}
    model = None
    torch.cuda.empty_cache()
    return line, extra


def run_ours(args, name, cfg, rank, world, local):
    from transvae import _lib
    _lib.require_device()
    C = Ctx(args, rank, world, local)
    if cfg["kind"] == "train":
        line, extra = run_train(C, name, cfg)
    else:
        line, extra = run_infer(C, name, cfg, args.steps, args.warmup)
    legs = world == 1 and not args.no_legs
    if legs and name == "train256":
        # secondary legs, N = 1 only: BASELINE configs[1] on our kernels, stock torch on the same GPU, the CPU reference
        il, ie = run_infer(C, "infer256", CONFIGS["infer256"], 5, 3)
        line["infer"] = {"config": "encode+decode b64", "value": il["value"], "e2e": il["e2e"]["value"] if il["e2e"] else None,
                         "model_frac": il["model_frac_of_peak"], "mtgemm_frac": il["roofline"]["frac"],
                         "attn_fwd_frac": il["roofline"]["attn_fwd_frac"]}
        extra["infer"] = ie
        line["gpu_stock_torch"] = stock_torch_leg(cfg, C.dev)
    if legs:
        rate, ms_cpu, cores, kind, sample = cpu_reference_rate(cfg, 1, 1)
        line["cpu_baseline"] = {"value": round(rate, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    else:
        line["cpu_baseline"] = None
    if rank != 0:
        return
    if args.breakdown_json:
        os.makedirs(os.path.dirname(os.path.abspath(args.breakdown_json)), exist_ok=True)
        with open(args.breakdown_json, "w") as fh:
            json.dump({"line": line, **extra}, fh, indent=1)
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        # CPU arm: rank 0 alone runs and prints; the other ranks exit 0 without work (no process group needed)
        run_reference(args, cfg, int(os.environ.get("RANK", "0")))
        return
    rank, world, local = dist_setup(args.gpus)
    try:
        run_ours(args, args.config, cfg, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
