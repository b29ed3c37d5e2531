# round 2, job f: full GPU test-suite (ILU, idrs_original, deterministic scatter, pins, total mesh) + default bench with all extras
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
timeout 1800 python bench.py --steps 3 --warmup 3 > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err
tail -n 15 gpurun_out/r2f_pytest.log | cut -c1-300
cut -c1-200 gpurun_out/r2f_bench.log | tail -n 1
tail -n 3 gpurun_out/r2f_bench.err
