#!/bin/bash
# Round-1 evidence capture on the GPU box (one GPU): launch list of the bench command, then ncu --set full captures of
# the dominant tensor kernel, both attention kernels, the wgrad kernel and representative HBM kernels at BASELINE-size
# shapes (tools/one_kernel.py).  Outputs land in gpurun_out/.
python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --no-e2e > gpurun_out/r1_final_plain.json 2> gpurun_out/r1_final_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_infer_b64_final.csv \
    python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --no-e2e > gpurun_out/ncu_ll.log 2>&1
cap() {  # name, kernel regex, one_kernel argument
  python tools/one_kernel.py $3 > /dev/null 2>&1 || { echo "plain run of $3 failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/r1_prof_$1 \
      python tools/one_kernel.py $3 > gpurun_out/ncu_$1.log 2>&1
}
cap mtgemm2_conv192 mtgemm2_kernel conv192
cap attn_fwd6 attn_fwd6_kernel attn_fwd
cap attn_bwd attn_bwd_kernel attn_bwd
cap wgrad192 mtwgrad_kernel wgrad192
cap gn_apply gn_apply_kernel gn_apply
cap gn_bwd_reduce gn_bwd_reduce_kernel gn_bwd
cap gn_bwd_apply gn_bwd_apply_kernel gn_bwd
ls -la gpurun_out/*.ncu-rep
