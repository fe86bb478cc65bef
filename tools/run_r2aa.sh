timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py tests/test_train_gpu.py tests/test_driver_gpu.py -x -q -m gpu > gpurun_out/r2aa_test.log 2>&1; tail -3 gpurun_out/r2aa_test.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-legs --no-e2e --breakdown-json gpurun_out/r2aa_bd.json > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; cut -c1-120 gpurun_out/r2aa_bench.json
timeout 400 python tools/at_kernel_sources.py large 32 > gpurun_out/r2aa_at_sources.txt 2>&1; head -8 gpurun_out/r2aa_at_sources.txt
