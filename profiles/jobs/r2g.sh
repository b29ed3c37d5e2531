# round 2, job g: full GPU test-suite after the residual fix + fused idrs; default bench; ncu launch list and full captures
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
MFB_KRYLOV_LEGACY=1 timeout 300 python -m pytest tests/test_krylov_gpu.py tests/test_golden_gpu.py -m gpu -q > gpurun_out/r2g_pytest_legacy.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest_legacy.log
timeout 1800 python bench.py --steps 3 --warmup 3 > gpurun_out/r2g_bench.log 2> gpurun_out/r2g_bench.err
bash profiles/ncu_job.sh r2g > gpurun_out/r2g_ncu.log 2>&1
tail -n 8 gpurun_out/r2g_pytest.log | cut -c1-300
tail -n 3 gpurun_out/r2g_pytest_legacy.log
cut -c1-200 gpurun_out/r2g_bench.log | tail -n 1
tail -n 3 gpurun_out/r2g_bench.err
