# round 2, job a: validate the fused Krylov path + multi-row SpMV; A/B against the round-1 launch structure
# usage (under gpurun): bash profiles/jobs/r2a.sh
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
MFB_SPMV=row timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_krylov_gpu.py -m gpu -q > gpurun_out/r2a_pytest_row.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_row.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_bench_new.log 2> gpurun_out/r2a_bench_new.err
MFB_SPMV=row timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_bench_newkrylov_rowspmv.log 2> gpurun_out/r2a_bench_newkrylov_rowspmv.err
MFB_KRYLOV_LEGACY=1 MFB_SPMV=row timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_bench_legacy.log 2> gpurun_out/r2a_bench_legacy.err
timeout 300 python bench.py --spmv-sweep > gpurun_out/r2a_sweep.log 2>&1
tail -3 gpurun_out/r2a_pytest.log gpurun_out/r2a_pytest_row.log
cat gpurun_out/r2a_bench_new.log | cut -c1-600
cat gpurun_out/r2a_sweep.log | tail -2
