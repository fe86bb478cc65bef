#!/bin/bash
# 8-GPU A/B of NCCL-registered gradient buffers (one box): overlapped buckets / one all-reduce after backward, each with the
# flat gradient buffer allocated by ncclMemAlloc + registered, against the plain allocation.
TAG=${1:-r2ae}
run() { # name, env, extra args...
  local name=$1; local ev=$2; shift; shift
  env $ev timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-legs --no-e2e "$@" > gpurun_out/${TAG}_8gpu_${name}.json 2> gpurun_out/${TAG}_8gpu_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
for l in open("gpurun_out/${TAG}_8gpu_${name}.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["value"], d["ms_per_step"], d["comm"], d["roofline"]["frac"], d["roofline"]["wgrad_frac"])
PY
  grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/${TAG}_8gpu_${name}.err | tail -3
}
run overlap_reg "TVAE_DDP_OVERLAP=1 TVAE_DDP_REGISTER=1"
run tail_reg "TVAE_DDP_OVERLAP=0 TVAE_DDP_REGISTER=1"
run overlap_plain "TVAE_DDP_OVERLAP=1 TVAE_DDP_REGISTER=0"
