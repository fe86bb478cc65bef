#!/bin/bash
# Evidence capture of round 2 (one GPU): the default bench line (all legs), the training launch list + census, and
# ncu --set full captures of the kernels that changed (each only after its command has run once without ncu).
T=${1:-r2r}
python bench.py --steps 3 --warmup 3 --breakdown-json gpurun_out/${T}_breakdown.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || { echo bench failed; tail -5 gpurun_out/${T}_bench.err; exit 1; }
python tools/train_probe.py large 32 > gpurun_out/${T}_train_probe.log 2>&1 || exit 2
TVAE_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${T}_launches_train_mb32.csv python tools/train_probe.py large 32 > gpurun_out/${T}_ncu_train.log 2>&1
python tools/ncu_census.py gpurun_out/${T}_launches_train_mb32.csv 60 > gpurun_out/${T}_census_train_mb32.txt
for k in attn_bwd wgrad192 token_norm_bwd gn_bwd; do
  python tools/one_kernel.py $k 16 > /dev/null 2>&1 || { echo "one_kernel $k failed"; continue; }
  case $k in
    attn_bwd) rx=attn_bwd_kernel;; wgrad192) rx=mtwgrad_h_kernel;; token_norm_bwd) rx=token_norm_bwd_reg_kernel;; gn_bwd) rx=gn_bwd_reduce_kernel;;
  esac
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/${T}_prof_$k python tools/one_kernel.py $k 16 > gpurun_out/${T}_ncu_$k.log 2>&1
done
ls -la gpurun_out/${T}_*
