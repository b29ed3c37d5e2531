# round 2, job c: SpMV sweep incl. TMA-ring and load-policy variants; idrs_original parity
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --spmv-sweep > gpurun_out/r2c_sweep.log 2>&1
timeout 300 python -m pytest tests/test_krylov_gpu.py -m gpu -q > gpurun_out/r2c_pytest_krylov.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_krylov.log
tail -n 2 gpurun_out/r2c_sweep.log; tail -n 3 gpurun_out/r2c_pytest_krylov.log
