# round 2, job e: deterministic scatter + everything since job b on one GPU; default bench with extras and cpu baseline
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2e_bench.log 2> gpurun_out/r2e_bench.err
MFB_DETERMINISTIC=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_bench_nondet.log 2> gpurun_out/r2e_bench_nondet.err
tail -n 4 gpurun_out/r2e_pytest.log
cut -c1-300 gpurun_out/r2e_bench.log | tail -n 1
tail -n 3 gpurun_out/r2e_bench.err
