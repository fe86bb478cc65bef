#!/bin/bash
# 8-GPU evidence run (one box): BASELINE configs[2] (DDP training, fp32 and bf16 gradient all-reduce), configs[3]
# (batch-sharded 512^2 / 1024^2 inference) and configs[4] (giant DDP).  Each command under its own timeout.
TAG=${1:-r2j}
run() { # name, extra args...
  local name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus 8 "$@" > gpurun_out/${TAG}_8gpu_${name}.json 2> gpurun_out/${TAG}_8gpu_${name}.err
  echo "== $name rc=$?"; tail -c 2600 gpurun_out/${TAG}_8gpu_${name}.json; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/${TAG}_8gpu_${name}.err | tail -3
}
run train256 --steps 4 --warmup 3 --breakdown-json gpurun_out/${TAG}_8gpu_train256_breakdown.json
run train256_bf16 --steps 4 --warmup 3 --grad-comm bf16
run extrap512 --config extrap512 --steps 4 --warmup 3
run extrap1024 --config extrap1024 --steps 3 --warmup 3
run giant --config giant --steps 2 --warmup 3
