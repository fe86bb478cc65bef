"""B200-native TransVAE training driver -- the caller of the hot path.

Drop-in for the reference's stage-1 recipe (train_2.py:276-405, train.py:579-620, train_working.py:305-436): bf16
compute with fp32 master weights, linear LR warm-up then constant, gradient clipping, fused AdamW(lr 1e-4,
betas (0.9, 0.95), wd 0) over flat buckets, gradient accumulation WITHOUT the reference's redundant per-micro-step
all-reduce, skip of non-finite steps (decided on the device, no host sync), checkpoints in the reference's dict
schema ({epoch, global_step, model_state_dict, optimizer_state_dict, scheduler_state_dict, args}) that either trainer
can resume from.  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        deepl-project_b200/train.py --variant large --resolution 256 --batch_size 32 --max_steps 1000 --output_dir out

Data is synthetic by default (seeded uniform images per rank; the reference's ImageNet / COCO loaders are out of
scope) or a tensor file of images (--data_tensor: [N, 3, H, W] uint8 or float in [0, 1]).  LPIPS / VF / GAN terms need
networks that are unavailable offline: their weights must be 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def parse_args(argv=None):
    p = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    # model (train.py:31-35)
    p.add_argument("--config", type=str, default=None, help="YAML with a `model:` section (reference configs/*.yaml)")
    p.add_argument("--variant", type=str, default="large", choices=["tiny", "base", "large", "huge", "giant"])
    p.add_argument("--compression_ratio", type=int, default=16, choices=[8, 16])
    p.add_argument("--latent_dim", type=int, default=32)
    # data (train.py:38-43)
    p.add_argument("--data_tensor", type=str, default=None, help=".pt file holding [N, 3, H, W] images; default: synthetic")
    p.add_argument("--resolution", type=int, default=256)
    p.add_argument("--batch_size", type=int, default=32, help="per-GPU micro-batch")
    # optimisation (train.py:46-49, train_2.py:60-70)
    p.add_argument("--num_epochs", type=int, default=1)
    p.add_argument("--steps_per_epoch", type=int, default=1000, help="synthetic data: optimiser steps per epoch")
    p.add_argument("--max_steps", type=int, default=None, help="stop after this many optimiser steps in total")
    p.add_argument("--learning_rate", type=float, default=1e-4)
    p.add_argument("--warmup_steps", type=int, default=1000)
    p.add_argument("--grad_clip", type=float, default=1.0)
    p.add_argument("--accumulation_steps", type=int, default=1)
    p.add_argument("--weight_decay", type=float, default=0.0)
    # loss (train.py:53-58)
    p.add_argument("--l1_weight", type=float, default=1.0)
    p.add_argument("--lpips_weight", type=float, default=0.0)
    p.add_argument("--kl_weight", type=float, default=1e-8)
    p.add_argument("--vf_weight", type=float, default=0.0)
    p.add_argument("--gan_weight", type=float, default=0.0)
    # checkpoints / logging (train.py:61-64)
    p.add_argument("--output_dir", type=str, required=True)
    p.add_argument("--checkpoint", type=str, default=None, help="resume from this checkpoint (ours or the reference's)")
    p.add_argument("--save_freq", type=int, default=5000, help="optimiser steps between checkpoints")
    p.add_argument("--log_freq", type=int, default=10)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--bucket_mb", type=int, default=64)
    p.add_argument("--grad_comm", type=str, default="fp32", choices=["fp32", "bf16"],
                   help="dtype of the gradient all-reduce (fp32 = DistributedDataParallel's; bf16 halves the payload)")
    return p.parse_args(argv)


def lr_at(step: int, base_lr: float, warmup_steps: int) -> float:
    """Learning rate of optimiser update `step` (0-based): the LambdaLR of train_2.py:266-274."""
    if warmup_steps > 0 and step < warmup_steps:
        return base_lr * step / max(1, warmup_steps)
    return base_lr


class SyntheticImages:
    """Seeded uniform [0, 1] images generated on the device; a different stream per rank, reproducible across resumes."""

    def __init__(self, batch: int, res: int, device, seed: int):
        self.shape, self.device, self.seed = (batch, 3, res, res), device, seed

    def batch(self, index: int) -> torch.Tensor:
        g = torch.Generator(device=self.device).manual_seed(self.seed * 1_000_003 + index)
        return torch.rand(self.shape, generator=g, device=self.device)


class TensorImages:
    def __init__(self, path: str, batch: int, rank: int, world: int, device):
        t = torch.load(path, map_location="cpu", weights_only=True)
        t = t["images"] if isinstance(t, dict) else t
        self.data = (t.float() / 255.0 if t.dtype == torch.uint8 else t.float())[rank::world].pin_memory()
        self.batch_size, self.device = batch, device
        if self.data.shape[0] < batch:
            raise ValueError(f"rank {rank} holds {self.data.shape[0]} images, fewer than one batch of {batch}")

    def batch(self, index: int) -> torch.Tensor:
        n = self.data.shape[0] // self.batch_size
        i = (index % n) * self.batch_size
        return self.data[i:i + self.batch_size].to(self.device, non_blocking=True)


def main(argv=None):
    args = parse_args(argv)
    if max(args.lpips_weight, args.vf_weight, args.gan_weight) > 0:
        raise NotImplementedError("LPIPS / VF / GAN terms need VGG / DINOv2 / discriminator weights (unavailable offline)")
    import transvae
    from transvae import _lib
    from transvae.trainer import Trainer
    _lib.require_device()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(args.seed)                          # identical initial replicas on every rank
    cfg = None
    if args.config:
        import yaml
        cfg = yaml.safe_load(open(args.config))["model"]
    with torch.device(dev):
        model = transvae.TransVAE(config=cfg, variant=args.variant, compression_ratio=args.compression_ratio,
                                  latent_dim=args.latent_dim, patched=True)
    loss_fn = transvae.TransVAELoss(l1_weight=args.l1_weight, lpips_weight=0.0, kl_weight=args.kl_weight, vf_weight=0.0,
                                    gan_weight=0.0)
    tr = Trainer(model, loss_fn, lr=args.learning_rate, betas=(0.9, 0.95), weight_decay=args.weight_decay,
                 grad_clip=args.grad_clip, accumulation_steps=args.accumulation_steps, bucket_bytes=args.bucket_mb << 20,
                 warmup_steps=args.warmup_steps, grad_comm=torch.bfloat16 if args.grad_comm == "bf16" else torch.float32)
    start_epoch = 0
    if args.checkpoint:
        ck = tr.load(args.checkpoint)
        start_epoch = int(ck.get("epoch", 0))
        if rank == 0:
            print(f"resumed from {args.checkpoint}: global_step {tr.opt.step_count}", flush=True)
    data = (TensorImages(args.data_tensor, args.batch_size, rank, world, dev) if args.data_tensor
            else SyntheticImages(args.batch_size, args.resolution, dev, args.seed * 977 + rank + 1))
    os.makedirs(args.output_dir, exist_ok=True)
    log = open(os.path.join(args.output_dir, "train_log.jsonl"), "a") if rank == 0 else None
    total = args.num_epochs * args.steps_per_epoch if args.max_steps is None else args.max_steps
    imgs_per_step = args.batch_size * args.accumulation_steps * world
    # `attempt` counts optimiser-step attempts on the host; the number of APPLIED updates lives on the device
    # (FusedAdamW.state[4]: a step with a non-finite gradient norm is skipped there and moves neither the bias correction
    # nor the warm-up, like the reference's `continue` ahead of optimizer.step() / scheduler.step(), train_2.py:329-338)
    # and is read back only when logging / saving, so the loop never synchronises with the device.
    attempt = applied0 = tr.opt.step_count
    t0 = time.perf_counter()
    while attempt < total:
        for a in range(args.accumulation_steps):
            out = tr.train_step(data.batch(attempt * args.accumulation_steps + a))
        attempt += 1
        if rank == 0 and (attempt % args.log_freq == 0 or attempt == total):
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            applied, skipped = tr.opt.step_count, tr.opt.skipped_steps
            rec = {"step": attempt, "applied_updates": applied, "skipped_updates": skipped,
                   "epoch": attempt // args.steps_per_epoch, "lr": lr_at(max(applied - 1, 0), args.learning_rate, args.warmup_steps),
                   "images_per_sec": imgs_per_step * (attempt - applied0) / dt, **{k: float(v) for k, v in out.items()}}
            print(json.dumps(rec), flush=True)
            log.write(json.dumps(rec) + "\n")
            log.flush()
        if attempt % args.save_freq == 0 or attempt == total:
            if rank == 0:
                path = os.path.join(args.output_dir, f"checkpoint_step{attempt}.pth")
                tr.save(path, epoch=max(start_epoch, attempt // args.steps_per_epoch), args=vars(args))
                print(f"Checkpoint saved to {path}", flush=True)
            if world > 1:
                dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
