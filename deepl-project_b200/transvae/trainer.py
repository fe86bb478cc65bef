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
    gloo, which is how the host logic is tested without GPUs)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20,
                 process_group: Optional[dist.ProcessGroup] = None):
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
            total += (p.numel() + 3) // 4 * 4            # keep every tensor 16-byte aligned
        self.numel = total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self._slices = {}
        with torch.no_grad():
            for p, o in zip(order, offs):
                n = p.numel()
                self.flat_p[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + n].view(p.shape)
                p.grad = self.flat_g[o:o + n].view(p.shape)
                # matrix-shaped weights: the wgrad kernel may accumulate straight into this slot (_autograd._wgrad_b)
                p._tvae_direct_grad = True
                self._slices[p] = (o, n)
        # contiguous buckets
        self.buckets: List[Tuple[int, int]] = []
        self._bucket_of = {}
        self._bucket_count: List[int] = []
        start, cnt, limit = 0, 0, max(1, bucket_bytes // 4)
        for p, o in zip(order, offs):
            n = (p.numel() + 3) // 4 * 4
            if cnt and (o + n - start) > limit:
                self.buckets.append((start, o))
                self._bucket_count.append(cnt)
                start, cnt = o, 0
            self._bucket_of[p] = len(self.buckets)
            cnt += 1
        self.buckets.append((start, total))
        self._bucket_count.append(cnt)
        self._pending = [0] * len(self.buckets)
        self._handles = []
        self.sync_grads = True          # False during gradient-accumulation micro-steps
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    # ------------------------------------------------------------------
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
            if self.sync_grads and self.world > 1:
                s, e = self.buckets[b]
                self._handles.append(dist.all_reduce(self.flat_g[s:e], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def wait(self) -> None:
        for h in self._handles:
            h.wait()
        self._handles.clear()

    def zero_grad(self) -> None:
        self.flat_g.zero_()
        self._pending = [0] * len(self.buckets)


class FusedAdamW:
    """AdamW over the flat buffers of ``GradBuckets`` with fused clipping / scaling / non-finite skip."""

    def __init__(self, buckets: GradBuckets, lr: float = 1e-4, betas: Tuple[float, float] = (0.9, 0.95), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: float = 1.0):
        self.b = buckets
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.m = torch.zeros_like(buckets.flat_p)
        self.v = torch.zeros_like(buckets.flat_p)
        self.ctrl = torch.zeros(4, dtype=torch.float32, device=buckets.flat_p.device)
        self.step_count = 0

    def step(self, grad_scale: float = 1.0, lr: Optional[float] = None) -> Tensor:
        """Applies one update; returns the (device) control block whose [0] is the sum of squared raw gradients."""
        from . import ops
        self.step_count += 1
        self.ctrl.zero_()
        ops.sumsq(self.b.flat_g, self.ctrl)
        self.ctrl[1] = self.max_norm if self.max_norm else 0.0
        self.ctrl[2] = grad_scale
        ops.adamw(self.b.flat_p, self.b.flat_g, self.m, self.v, self.ctrl, self.lr if lr is None else lr, self.betas,
                  self.eps, self.wd, self.step_count)
        ops.WEIGHT_EPOCH += 1           # the kernel rewrote flat_p: packed operands cached for this step are stale
        return self.ctrl

    # ---- checkpoint interchange --------------------------------------------------------------------------------
    # The reference saves ``torch.optim.AdamW.state_dict()`` (train.py:753-769, train_2.py:245-260).  The same schema
    # is produced / accepted here -- state[i] = {step, exp_avg, exp_avg_sq} with i the index of the parameter in
    # ``model.parameters()`` order (trainable ones only), one param group -- so a run can move between the reference's
    # trainer and this one in either direction.
    def state_dict(self) -> dict:
        state = {}
        for i, p in enumerate(self.b.params):
            o, n = self.b._slices[p]
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.m[o:o + n].view(p.shape).clone(),
                        "exp_avg_sq": self.v[o:o + n].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
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


class Trainer:
    """model(images) -> TransVAELoss -> backward (overlapped all-reduce) -> clip -> AdamW, one call per micro-batch."""

    def __init__(self, model: torch.nn.Module, loss_fn: torch.nn.Module, lr: float = 1e-4,
                 betas: Tuple[float, float] = (0.9, 0.95), weight_decay: float = 0.0, grad_clip: float = 1.0,
                 accumulation_steps: int = 1, bucket_bytes: int = 64 << 20, warmup_steps: int = 0,
                 process_group: Optional[dist.ProcessGroup] = None):
        self.model, self.loss_fn = model, loss_fn
        self.buckets = GradBuckets(model.parameters(), bucket_bytes, process_group)
        self.opt = FusedAdamW(self.buckets, lr, betas, 1e-8, weight_decay, grad_clip)
        self.accum = max(1, accumulation_steps)
        self.warmup_steps = warmup_steps
        self._micro = 0
        self.world = self.buckets.world

    def _lr(self) -> float:
        # linear warm-up then constant (train_2.py:266-274: LambdaLR with step / warmup, so update k uses k / warmup)
        if self.warmup_steps > 0:
            return self.opt.lr * min(1.0, self.opt.step_count / self.warmup_steps)
        return self.opt.lr

    def train_step(self, images: Tensor, eps: Optional[Tensor] = None) -> dict:
        """One micro-batch.  Returns the loss dict (device tensors; nothing is synchronised with the host)."""
        self.model.train()
        last = (self._micro + 1) % self.accum == 0
        self.buckets.sync_grads = last
        recon, mu, logvar = self.model(images, eps=eps)
        losses = self.loss_fn(recon, images, mu, logvar)
        losses["total"].backward()
        self._micro += 1
        if last:
            self.buckets.wait()
            self.opt.step(grad_scale=1.0 / (self.world * self.accum), lr=self._lr())
            self.buckets.zero_grad()
        return {k: v.detach() for k, v in losses.items()}

    # checkpoint schema of the reference (train.py:753-769, train_2.py:245-260)
    def state_dict(self, epoch: int = 0, args: Optional[dict] = None) -> dict:
        return {"epoch": epoch, "global_step": self.opt.step_count, "model_state_dict": self.model.state_dict(),
                "optimizer_state_dict": self.opt.state_dict(),
                "scheduler_state_dict": {"last_epoch": self.opt.step_count, "warmup_steps": self.warmup_steps},
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

    def save(self, path: str, epoch: int = 0, args: Optional[dict] = None) -> None:
        torch.save(self.state_dict(epoch, args), path)

    def load(self, path: str) -> dict:
        sd = torch.load(path, map_location=self.buckets.flat_p.device, weights_only=False)
        self.load_state_dict(sd)
        return sd
