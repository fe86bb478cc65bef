set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py tests/test_train_gpu.py tests/test_model_gpu.py -x -q -m gpu > gpurun_out/r2y_test.log 2>&1; tail -3 gpurun_out/r2y_test.log
(
for v in 1 0; do for k in lin384_dgelu lin384_res; do echo "== $k TVAE_EPI_PIPE_RES=$v"; TVAE_EPI_PIPE_RES=$v timeout 100 python tools/time_kernel.py $k 32; done; done
) > gpurun_out/r2y_micro.log 2>&1
grep -v "^+" gpurun_out/r2y_micro.log
for v in 1 0; do TVAE_EPI_PIPE_RES=$v timeout 300 python bench.py --steps 2 --warmup 3 --no-legs --no-e2e --breakdown-json gpurun_out/r2y_bd_pipe$v.json > gpurun_out/r2y_bench_pipe$v.json 2> gpurun_out/r2y_bench_pipe$v.err; cut -c1-120 gpurun_out/r2y_bench_pipe$v.json; done
