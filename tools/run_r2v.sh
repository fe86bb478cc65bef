set -x
timeout 600 python -m pytest tests/test_backward_gpu.py tests/test_train_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/r2v_test.log 2>&1; tail -3 gpurun_out/r2v_test.log
(
for v in 1 0; do echo "== TVAE_ATTN_BWD_CHAINS=$v"; TVAE_ATTN_BWD_CHAINS=$v timeout 200 python tools/attn_bench.py 32 --bwd; done
echo "== unordered"; TVAE_ATTN_BWD_ORDERED=0 timeout 200 python tools/attn_bench.py 32 --bwd
for k in dgrad192 dgrad192gnb; do for h in 1 0; do echo "== $k TVAE_HALO=$h"; TVAE_HALO=$h timeout 100 python tools/time_kernel.py $k 32; done; done
for k in wgrad_lin384 wgrad_lin384t wgrad_lin768 lin384_dgelu lin384_dual qkv384; do timeout 100 python tools/time_kernel.py $k 32; done
) > gpurun_out/r2v_micro.log 2>&1
cat gpurun_out/r2v_micro.log
