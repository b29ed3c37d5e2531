# round 2, job w1: single-wave grids for the light kernels -- Krylov / parity tests, then the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_krylov_gpu.py tests/test_parity_gpu.py tests/test_golden_gpu.py -m gpu -q > gpurun_out/r2w1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w1_pytest.log
tail -n 3 gpurun_out/r2w1_pytest.log | cut -c1-300
timeout 1800 python bench.py > gpurun_out/r2w1_bench.log 2> gpurun_out/r2w1_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w1_bench.log").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "solve_ms", "spmv_ms", "non_spmv_ms_per_spmv", "launches_per_spmv")}, d["config"]["krylov_iterations_per_step"], d["clocks"])
print(d["e2e"])
for k, e in d["extra"].items():
    print(k, e.get("ms_per_step"), e.get("krylov_iterations"), e.get("non_spmv_ms_per_spmv"), e.get("solve_ms"))
PY
tail -n 3 gpurun_out/r2w1_bench.err | cut -c1-300
