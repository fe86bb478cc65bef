#!/bin/bash
# 8-GPU A/B of the gradient all-reduce placement (one box): overlapped buckets (default) vs ONE all-reduce after backward
# (TVAE_DDP_OVERLAP=0), fp32 and bf16 payload.
TAG=${1:-r2w}
run() { # name, env, extra args...
  local name=$1; local ev=$2; shift; shift
  env $ev timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-legs --no-e2e "$@" > gpurun_out/${TAG}_8gpu_${name}.json 2> gpurun_out/${TAG}_8gpu_${name}.err
  echo "== $name rc=$?"; tail -c 1800 gpurun_out/${TAG}_8gpu_${name}.json; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/${TAG}_8gpu_${name}.err | tail -3
}
run overlap_fp32 TVAE_DDP_OVERLAP=1
run tail_fp32 TVAE_DDP_OVERLAP=0
run tail_bf16 TVAE_DDP_OVERLAP=0 --grad-comm bf16
