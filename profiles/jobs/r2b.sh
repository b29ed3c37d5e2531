# round 2, job b: TMA-ring SpMV + restructured fused vector kernel + pins + sparse-ids getter
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
MFB_SPMV=tma timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_krylov_gpu.py -m gpu -q > gpurun_out/r2b_pytest_tma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_tma.log
timeout 300 python bench.py --spmv-sweep > gpurun_out/r2b_sweep.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_mr.log 2> gpurun_out/r2b_bench_mr.err
MFB_SPMV=row timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_row.log 2> gpurun_out/r2b_bench_row.err
MFB_SPMV=tma timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_tma.log 2> gpurun_out/r2b_bench_tma.err
tail -n 3 gpurun_out/r2b_pytest.log; tail -n 3 gpurun_out/r2b_pytest_tma.log
tail -n 2 gpurun_out/r2b_sweep.log
