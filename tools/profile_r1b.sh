#!/bin/bash
# Evidence capture for the second half of round 1 (one GPU): launch lists of the inference bench command and of a
# training run, and one ncu --set full capture of the dominant kernel with the fused GroupNorm statistics.
python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --no-e2e > gpurun_out/r1b_plain.json 2> gpurun_out/r1b_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_launches_infer_b64.csv \
    python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --no-e2e > gpurun_out/r1b_ncu_ll.log 2>&1
python tools/train_probe.py large 32 > gpurun_out/r1b_train_probe.log 2>&1 || { echo "train_probe failed"; tail -5 gpurun_out/r1b_train_probe.log; }
TVAE_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r1b_launches_train_mb32.csv python tools/train_probe.py large 32 > gpurun_out/r1b_ncu_train.log 2>&1
python tools/one_kernel.py conv192gn > /dev/null 2>&1 || exit 2
ncu --set full --clock-control none --import-source on -k regex:mtgemm2_kernel -s 3 -c 1 -f -o gpurun_out/r1b_prof_mtgemm2_conv192gn \
    python tools/one_kernel.py conv192gn > gpurun_out/r1b_ncu_conv192gn.log 2>&1
ls -la gpurun_out/r1b_*
