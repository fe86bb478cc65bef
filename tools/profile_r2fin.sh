#!/bin/bash
# Final evidence capture of round 2 (one GPU): GPU tests, smoke, the default bench line (all legs) + reference arm, the
# training launch list + census, and ncu --set full captures of the kernels that changed in the third session (each only
# after its command has run once without ncu).
T=${1:-r2fin}
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
python bench.py --steps 3 --warmup 3 --breakdown-json gpurun_out/${T}_breakdown.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || { echo bench failed; tail -5 gpurun_out/${T}_bench.err; }
cat gpurun_out/${T}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; cut -c1-200 gpurun_out/${T}_bench_reference_arm.json
python tools/train_probe.py large 32 > gpurun_out/${T}_train_probe.log 2>&1
TVAE_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${T}_launches_train_mb32.csv python tools/train_probe.py large 32 > gpurun_out/${T}_ncu_train.log 2>&1
python tools/ncu_census.py gpurun_out/${T}_launches_train_mb32.csv 60 > gpurun_out/${T}_census_train_mb32.txt
for k in dgrad192gnb lin384_dgelu wgrad_lin384 attn_bwd; do
  python tools/one_kernel.py $k 16 > /dev/null 2>&1 || { echo "one_kernel $k failed"; continue; }
  case $k in
    attn_bwd) rx=attn_bwd_kernel;; wgrad_lin384) rx=mtwgrad;; *) rx=mtgemm2_kernel;;
  esac
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/${T}_prof_$k python tools/one_kernel.py $k 16 > gpurun_out/${T}_ncu_$k.log 2>&1
done
ls -la gpurun_out/${T}_* | head -30
