#!/bin/bash
# ncu --set full captures of the token-GEMM shapes at the stage-1 width (long pixel axis, K = 384): weight gradient, the
# GELU-gradient epilogue, the dual-output epilogue and the QKV / RoPE epilogue.  Each after a plain run of the same command.
T=${1:-r2x}
for k in wgrad_lin384 lin384_dgelu lin384_dual qkv384; do
  python tools/time_kernel.py $k 32 > gpurun_out/${T}_time_$k.log 2>&1 || { echo "time_kernel $k failed"; continue; }
  cat gpurun_out/${T}_time_$k.log | tail -1
  case $k in wgrad_lin384) rx='mtwgrad';; *) rx='mtgemm2_kernel';; esac
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/${T}_prof_$k python tools/one_kernel.py $k 32 > gpurun_out/${T}_ncu_$k.log 2>&1
done
ls -la gpurun_out/${T}_*
