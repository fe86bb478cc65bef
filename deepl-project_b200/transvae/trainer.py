"""Data-parallel training step for TransVAE on B200: the B200-native counterpart of the reference's training loop
(train.py:579-620, train_working.py:331-400): forward -> L1+KL loss -> backward -> gradient all-reduce -> clip -> AdamW.

* one process per GPU (torchrun), batch sharded across ranks, replicas of all parameters;
* ``GradBuckets``: parameters and gradients live in flat fp32 buffers; gradients are grouped into contiguous buckets in
  reverse execution order (decoder.conv_out ... encoder.conv_in) and each bucket's NCCL all-reduce is launched from a
  post-accumulate-grad hook the moment its last gradient lands, so communication overlaps the rest of backward
  (what DistributedDataParallel does for the reference, train.py:672-674, minus its flatten copies; with gradient
  accumulation the all-reduce only fires on the last micro-step -- the reference has no no_sync(), train.py:599-603);
* ``FusedAdamW``: one ``tvae_sumsq`` + one ``tvae_adamw`` launch over the flat buffers; gradient clipping, the 1/world
  (and 1/accum) scaling and the skip-on-non-finite rule (train_2.py:329-338) are folded into the AdamW kernel through a
  4-float device control block, so the step needs no host synchronisation.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


class GradBuckets:
    """Flat fp32 parameter / gradient storage with bucketed, overlapped all-reduce.  Pure torch (works on CPU with
    gloo, which is how the host logic is tested without GPUs).

    ``grad_comm``: dtype of the all-reduce payload.  ``torch.float32`` (default) is DistributedDataParallel's behaviour
    (train.py:672-674).  ``torch.bfloat16`` halves the NVLink payload: a bucket is cast into a bf16 mirror
    (``tvae_cast_f32_bf16``) the moment it is complete, the mirror is all-reduced and the optimizer reads it -- the
    rounding of every gradient to bf16 (2^-9 relative) is a stated deviation from the reference.

    ``overlap``: True (default; ``TVAE_DDP_OVERLAP=0`` turns it off) launches a bucket's all-reduce from the hook of its
    last gradient, so it runs under the rest of backward; False sends ONE all-reduce over the whole flat buffer at the end
    of backward (``wait()``) -- nothing competes with the persistent GEMM grids for SMs, the transfer is exposed."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20,
                 process_group: Optional[dist.ProcessGroup] = None, grad_comm: torch.dtype = torch.float32,
                 overlap: Optional[bool] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        from . import _autograd, ops
        _autograd.clear_pack_cache()
        ops.WEIGHT_EPOCH += 1
        dev = self.params[0].device
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        # reverse registration order ~ the order in which backward produces gradients
        order = list(reversed(self.params))
        offs, total = [], 0
        for p in order:
            offs.append(total)
            total += (p.numel() + 7) // 8 * 8            # every tensor 32-byte aligned (16 bytes in the bf16 mirror)
        self.numel = total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self._pools = []                # NCCL-registered memory pools of the communication buffers (kept alive)
        self.registered = False
        self.flat_g = self._comm_zeros(total, torch.float32, dev)
        if grad_comm not in (torch.float32, torch.bfloat16):
            raise ValueError("grad_comm must be torch.float32 or torch.bfloat16")
        self.grad_comm = grad_comm
        # bf16 mirror of the gradients (communication + optimizer input) -- only with more than one rank
        self.comm_g = self._comm_zeros(total, torch.bfloat16, dev) if (grad_comm == torch.bfloat16 and self.world > 1) else None
        self._slices = {}
        with torch.no_grad():
            for p, o in zip(order, offs):
                n = p.numel()
                self.flat_p[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + n].view(p.shape)
                p.grad = self.flat_g[o:o + n].view(p.shape)
                # the wgrad kernels may accumulate straight into this slot (_autograd._wgrad_b / GradSink)
                p._tvae_direct_grad = True
                self._slices[p] = (o, n)
            if self.world > 1:
                # identical replicas on every rank (DistributedDataParallel broadcasts rank 0's parameters too)
                dist.broadcast(self.flat_p, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
        # contiguous buckets
        self.buckets: List[Tuple[int, int]] = []
        self._bucket_of = {}
        self._bucket_count: List[int] = []
        start, cnt, limit = 0, 0, max(1, bucket_bytes // 4)
        for p, o in zip(order, offs):
            n = (p.numel() + 7) // 8 * 8
            if cnt and (o + n - start) > limit:
                self.buckets.append((start, o))
                self._bucket_count.append(cnt)
                start, cnt = o, 0
            self._bucket_of[p] = len(self.buckets)
            cnt += 1
        self.buckets.append((start, total))
        self._bucket_count.append(cnt)
        self._pending = [0] * len(self.buckets)
        self._reduced = [False] * len(self.buckets)
        self._handles = []
        self.sync_grads = True          # False during gradient-accumulation micro-steps
        if overlap is None:
            import os
            overlap = os.environ.get("TVAE_DDP_OVERLAP", "1") != "0"
        self.overlap = bool(overlap)
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    # ------------------------------------------------------------------
    def _comm_zeros(self, numel: int, dtype: torch.dtype, dev) -> Tensor:
        """Zero-filled communication buffer.  With NCCL and more than one rank it comes from the communicator's own
        allocator (ncclMemAlloc) and is registered with it (ncclCommRegister, through ``register_mem_pool``): the all-reduce
        then reads / reduces user memory in place (NVLS over the NVSwitch) instead of staging through NCCL's bounce
        buffers -- 8 GPUs, same box: 978 -> 1014 img/s with overlapped buckets (profiles/r2ae_8gpu_*).
        ``TVAE_DDP_REGISTER=0`` or any failure of that route: a plain torch allocation."""
        import os
        want = (self.world > 1 and torch.device(dev).type == "cuda" and os.environ.get("TVAE_DDP_REGISTER", "1") != "0"
                and dist.get_backend(self.pg) == "nccl")
        if want:
            try:
                backend = (self.pg if self.pg is not None else dist.group.WORLD)._get_backend(torch.device(dev))
                pool = torch.cuda.MemPool(backend.mem_allocator)
                with torch.cuda.use_mem_pool(pool):
                    t = torch.zeros(numel, dtype=dtype, device=dev)
                backend.register_mem_pool(pool)
                self._pools.append((backend, pool))
                self.registered = True
                return t
            except Exception as e:          # noqa: BLE001 -- older torch / NCCL without the allocator: fall back, say so once
                import warnings
                warnings.warn(f"NCCL buffer registration unavailable ({type(e).__name__}: {e}); plain allocation")
        return torch.zeros(numel, dtype=dtype, device=dev)

    def grads(self) -> Tensor:
        """The flat gradient buffer the optimizer reads: the all-reduced bf16 mirror, or the fp32 buffer."""
        return self.comm_g if self.comm_g is not None else self.flat_g

    def _reduce_bucket(self, b: int) -> None:
        s, e = self.buckets[b]
        self._reduced[b] = True
        if self.world <= 1:
            return
        if self.comm_g is not None:
            from . import ops
            ops.cast_f32_bf16(self.flat_g[s:e], self.comm_g[s:e])
            buf = self.comm_g[s:e]
        else:
            buf = self.flat_g[s:e]
        self._handles.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def _hook(self, p: torch.nn.Parameter) -> None:
        o, n = self._slices[p]
        if p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
            # autograd replaced .grad (e.g. after set_to_none): fold it back into the flat buffer
            with torch.no_grad():
                self.flat_g[o:o + n].add_(p.grad.reshape(-1))
                p.grad = self.flat_g[o:o + n].view(p.shape)
        b = self._bucket_of[p]
        self._pending[b] += 1
        if self._pending[b] == self._bucket_count[b]:
            self._pending[b] = 0
            if self.sync_grads and self.overlap:
                self._reduce_bucket(b)

    def _reduce_all(self) -> None:
        """One all-reduce over the whole flat buffer (``overlap=False``)."""
        self._reduced = [True] * len(self.buckets)
        if self.world <= 1:
            return
        if self.comm_g is not None:
            from . import ops
            ops.cast_f32_bf16(self.flat_g, self.comm_g)
        self._handles.append(dist.all_reduce(self.grads(), op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def wait(self) -> None:
        """End of backward.  Buckets that never completed -- a parameter without a gradient this step (unused branch,
        frozen after construction) -- are reduced now, so ranks cannot diverge silently."""
        if self.sync_grads and not self.overlap and not any(self._reduced):
            self._reduce_all()
        if self.sync_grads:
            for b in range(len(self.buckets)):
                if not self._reduced[b]:
                    self._reduce_bucket(b)
        for h in self._handles:
            h.wait()
        self._handles.clear()
        self._pending = [0] * len(self.buckets)
        self._reduced = [False] * len(self.buckets)

    def zero_grad(self) -> None:
        self.flat_g.zero_()
        self._pending = [0] * len(self.buckets)
        self._reduced = [False] * len(self.buckets)


class FusedAdamW:
    """AdamW over the flat buffers of ``GradBuckets`` with fused clipping / scaling / warm-up / non-finite skip.  The
    update counter lives on the device (``state[4]``): a skipped step advances neither the bias correction nor the
    warm-up, like the reference's ``continue`` ahead of optimizer.step() / scheduler.step() (train_2.py:329-338)."""

    def __init__(self, buckets: GradBuckets, lr: float = 1e-4, betas: Tuple[float, float] = (0.9, 0.95), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: float = 1.0, warmup_steps: int = 0):
        self.b = buckets
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.warmup_steps = warmup_steps
        self.m = torch.zeros_like(buckets.flat_p)
        self.v = torch.zeros_like(buckets.flat_p)
        dev = buckets.flat_p.device
        self.state = torch.zeros(8, dtype=torch.float32, device=dev)
        self.partials = torch.zeros(1024, dtype=torch.float64, device=dev)

    # number of APPLIED updates; reading it synchronises with the device (logging / checkpoints only)
    @property
    def step_count(self) -> int:
        return int(self.state[4].item())

    @step_count.setter
    def step_count(self, n: int) -> None:
        self.state[4] = float(n)

    @property
    def skipped_steps(self) -> int:
        return int(self.state[5].item())

    @property
    def ctrl(self) -> Tensor:
        """[0] = sum of squared (unscaled) gradients of the last step, [1] = its clip factor (device tensor)."""
        return self.state

    def current_lr(self) -> float:
        """Learning rate of the next update (LambdaLR of train_2.py:266-274: update k uses lr * min(1, k / warmup))."""
        k = self.step_count
        return self.lr * min(1.0, k / self.warmup_steps) if self.warmup_steps > 0 else self.lr

    def step(self, grad_scale: float = 1.0) -> Tensor:
        """One update attempt; returns the device state block.  Nothing is synchronised with the host."""
        from . import ops
        g = self.b.grads()
        ops.grad_sumsq(g, self.partials)
        ops.adamw_step(self.b.flat_p, g, self.m, self.v, self.partials, self.state, self.lr, self.warmup_steps, self.betas,
                       self.eps, self.wd, self.max_norm, grad_scale)
        ops.WEIGHT_EPOCH += 1           # the kernel rewrote flat_p: packed operands cached for this step are stale
        return self.state

    # ---- checkpoint interchange --------------------------------------------------------------------------------
    # The reference saves ``torch.optim.AdamW.state_dict()`` (train.py:753-769, train_2.py:245-260).  The same schema
    # is produced / accepted here -- state[i] = {step, exp_avg, exp_avg_sq} with i the index of the parameter in
    # ``model.parameters()`` order (trainable ones only), one param group -- so a run can move between the reference's
    # trainer and this one in either direction.
    def state_dict(self) -> dict:
        state = {}
        steps = self.step_count
        for i, p in enumerate(self.b.params):
            o, n = self.b._slices[p]
            state[i] = {"step": torch.tensor(float(steps)),
                        "exp_avg": self.m[o:o + n].view(p.shape).clone(),
                        "exp_avg_sq": self.v[o:o + n].view(p.shape).clone()}
        # what torch.optim.AdamW under a LambdaLR holds: `initial_lr` = the base rate, `lr` = the scheduled rate
        group = {"lr": self.current_lr(), "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": True, "initial_lr": self.lr,
                 "params": list(range(len(self.b.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        if "state" not in sd:                       # flat layout written by early versions of this trainer
            self.m.copy_(sd["m"])
            self.v.copy_(sd["v"])
            self.step_count = int(sd["step"])
            return
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(self.b.params):
            raise ValueError(f"optimizer state has {len(ids)} parameters, the model has {len(self.b.params)} trainable ones")
        steps = set()
        with torch.no_grad():
            for pid, p in zip(ids, self.b.params):
                st = sd["state"].get(pid)
                o, n = self.b._slices[p]
                if st is None:                      # parameter never updated: moments stay zero
                    self.m[o:o + n].zero_()
                    self.v[o:o + n].zero_()
                    continue
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state {pid}: shape {tuple(st['exp_avg'].shape)} != {tuple(p.shape)}")
                self.m[o:o + n].copy_(st["exp_avg"].reshape(-1))
                self.v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}); the fused optimiser keeps one")
        self.step_count = steps.pop() if steps else 0
        g0 = groups[0]
        self.lr = float(g0.get("initial_lr", g0.get("lr", self.lr)))      # LambdaLR keeps the base rate in initial_lr
        self.betas = tuple(g0.get("betas", self.betas))
        self.eps = float(g0.get("eps", self.eps))
        self.wd = float(g0.get("weight_decay", self.wd))


def lambda_lr_state(base_lr: float, warmup_steps: int, steps: int) -> dict:
    """``torch.optim.lr_scheduler.LambdaLR.state_dict()`` of the reference's warm-up schedule (train_2.py:266-274) after
    ``steps`` scheduler steps -- produced by a real LambdaLR on a one-element CPU optimizer, so that the reference's
    ``scheduler.load_state_dict(ckpt["scheduler_state_dict"])`` (train_2.py:489-490) accepts it on this torch version."""
    opt = torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=base_lr)
    fn = (lambda k: min(1.0, k / warmup_steps)) if warmup_steps > 0 else (lambda k: 1.0)
    sch = torch.optim.lr_scheduler.LambdaLR(opt, fn)
    sch.last_epoch = int(steps)
    sch._step_count = int(steps) + 1
    sch._last_lr = [base_lr * fn(int(steps))]
    return sch.state_dict()


class Trainer:
    """model(images) -> TransVAELoss -> backward (overlapped all-reduce) -> clip -> AdamW, one call per micro-batch."""

    def __init__(self, model: torch.nn.Module, loss_fn: torch.nn.Module, lr: float = 1e-4,
                 betas: Tuple[float, float] = (0.9, 0.95), weight_decay: float = 0.0, grad_clip: float = 1.0,
                 accumulation_steps: int = 1, bucket_bytes: int = 64 << 20, warmup_steps: int = 0,
                 process_group: Optional[dist.ProcessGroup] = None, grad_comm: torch.dtype = torch.float32):
        self.model, self.loss_fn = model, loss_fn
        self.buckets = GradBuckets(model.parameters(), bucket_bytes, process_group, grad_comm)
        self.opt = FusedAdamW(self.buckets, lr, betas, 1e-8, weight_decay, grad_clip, warmup_steps)
        self.accum = max(1, accumulation_steps)
        self.warmup_steps = warmup_steps
        self._micro = 0
        self.world = self.buckets.world
        self._h2d: Optional[torch.cuda.Stream] = None
        self._staged = None             # (host tensor, device copy, ready event) of the prefetched micro-batch
        self._loss_host: Optional[Tensor] = None
        # optional device-side timing of the step tail (bench.py): a list that receives, per optimiser step, the events
        # (end of backward, gradients all-reduced, optimiser + zero_grad done)
        self.timing: Optional[list] = None

    def train_step(self, images: Tensor, eps: Optional[Tensor] = None) -> dict:
        """One micro-batch.  Returns the loss dict (device tensors; nothing is synchronised with the host)."""
        self.model.train()
        last = (self._micro + 1) % self.accum == 0
        self.buckets.sync_grads = last
        _set_comm_active(last and self.world > 1)
        try:
            recon, mu, logvar = self.model(images, eps=eps)
            losses = self.loss_fn(recon, images, mu, logvar)
            losses["total"].backward()
            from . import _autograd
            _autograd.GRAD_SINK.flush()
        finally:
            _set_comm_active(False)
        self._micro += 1
        if last:
            ev = None
            if self.timing is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
            self.buckets.wait()
            if ev is not None:
                ev[1].record()
            self.opt.step(grad_scale=1.0 / (self.world * self.accum))
            self.buckets.zero_grad()
            if ev is not None:
                ev[2].record()
                self.timing.append(ev)
        return {k: v.detach() for k, v in losses.items()}

    # ---- host-to-host step: the reference's `images.to(device, non_blocking=True)` ... `loss.item()` loop ---------
    def _stage(self, x_host: Tensor) -> None:
        if not x_host.is_pinned():
            raise ValueError("host batches must be pinned (torch.Tensor.pin_memory) for asynchronous copies")
        dev = self.buckets.flat_p.device
        if self._h2d is None:
            self._h2d = torch.cuda.Stream(dev)
        with torch.cuda.stream(self._h2d):
            xd = x_host.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._h2d)
        self._staged = (x_host, xd, ev)

    def train_step_host(self, images_host: Tensor, next_images_host: Optional[Tensor] = None,
                        eps: Optional[Tensor] = None) -> Tensor:
        """``train_step`` on a PINNED host micro-batch (train.py:586 ``images.to(device, non_blocking=True)``): the upload
        of ``next_images_host`` runs on a copy stream under this step's kernels, and the loss terms come back into a
        pinned host tensor [6] = (l1, lpips, kl, vf, gan, total) (train.py:623 ``loss.item()``) with an asynchronous
        copy -- read it after ``torch.cuda.current_stream().synchronize()``."""
        dev = self.buckets.flat_p.device
        cur = torch.cuda.current_stream(dev)
        if self._staged is None or self._staged[0] is not images_host:
            self._stage(images_host)
        _, xd, ready = self._staged
        self._staged = None
        if next_images_host is not None:
            self._stage(next_images_host)
        cur.wait_event(ready)
        out = self.train_step(xd, eps=eps)
        xd.record_stream(cur)
        if self._loss_host is None:
            self._loss_host = torch.empty(6, dtype=torch.float32).pin_memory()
        vec = torch.stack([out[k].float().reshape(()) for k in ("l1", "lpips", "kl", "vf", "gan", "total")])
        self._loss_host.copy_(vec, non_blocking=True)
        return self._loss_host

    # checkpoint schema of the reference (train.py:753-769, train_2.py:245-260)
    def state_dict(self, epoch: int = 0, args: Optional[dict] = None) -> dict:
        steps = self.opt.step_count
        return {"epoch": epoch, "global_step": steps, "model_state_dict": self.model.state_dict(),
                "optimizer_state_dict": self.opt.state_dict(),
                "scheduler_state_dict": lambda_lr_state(self.opt.lr, self.warmup_steps, steps),
                "args": dict(args or {})}

    def load_state_dict(self, sd: dict) -> None:
        """Accepts checkpoints written by this trainer or by the reference's train.py / train_2.py."""
        self.model.load_state_dict(sd["model_state_dict"])
        # load_state_dict copies into the existing (flat-buffer backed) parameter storage, so the views stay valid
        from . import ops
        ops.WEIGHT_EPOCH += 1           # packed operands cached from the old weights are stale
        self.opt.load_state_dict(sd["optimizer_state_dict"])
        if "global_step" in sd and not sd["optimizer_state_dict"].get("state"):
            self.opt.step_count = int(sd["global_step"])
        sch = sd.get("scheduler_state_dict")
        if isinstance(sch, dict) and sch.get("base_lrs"):
            self.opt.lr = float(sch["base_lrs"][0])       # the scheduled `lr` of the param group is not the base rate

    def save(self, path: str, epoch: int = 0, args: Optional[dict] = None) -> None:
        torch.save(self.state_dict(epoch, args), path)

    def load(self, path: str) -> dict:
        sd = torch.load(path, map_location=self.buckets.flat_p.device, weights_only=False)
        self.load_state_dict(sd)
        return sd


def _set_comm_active(on: bool) -> None:
    """While gradient buckets are being all-reduced under the backward pass, the persistent GEMM kernels leave a few
    SMs to NCCL's channels (``tvae_set_reserved_sms``): a one-CTA-per-SM kernel whose grid exceeds the free SMs runs a
    nearly empty second wave -- up to 2x its time."""
    from . import ops
    ops.set_comm_active(on)
