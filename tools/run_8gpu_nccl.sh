#!/bin/bash
# 8-GPU A/B: CTAs NCCL may use for the gradient all-reduce vs SMs our persistent grids leave free while it runs
run() { # name, env..., --
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e > gpurun_out/r2k_8gpu_${name}.json 2> gpurun_out/r2k_8gpu_${name}.err
  echo "== $name rc=$?"; python - <<P
import json
d=json.loads(open("gpurun_out/r2k_8gpu_${name}.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["wgrad_frac"], d["comm"], d["clocks"]["sm_mhz"])
P
}
run cta4_res4 NCCL_MAX_CTAS=4 NCCL_MIN_CTAS=4 TVAE_COMM_RESERVED_SMS=4
run cta8_res8 NCCL_MAX_CTAS=8 NCCL_MIN_CTAS=8 TVAE_COMM_RESERVED_SMS=8
run cta4_res0 NCCL_MAX_CTAS=4 NCCL_MIN_CTAS=4 TVAE_COMM_RESERVED_SMS=0
run cta2_res2_bf16 NCCL_MAX_CTAS=2 NCCL_MIN_CTAS=2 TVAE_COMM_RESERVED_SMS=2
