for f in 4 8 14 24; do echo "== F=$f"; TVAE_WGRAD_SPLIT_F=$f timeout 200 python tools/train_probe.py large 32 2>&1 | grep -E "micro-batch|wgrad" ; done > gpurun_out/r2ac_split_f_sweep.log 2>&1
cat gpurun_out/r2ac_split_f_sweep.log
