# round 2: sanity of the rebuilt library (committed sources): Krylov tests, one timed step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_krylov_gpu.py -m gpu -q > gpurun_out/r2last_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2last_pytest.log
tail -n 2 gpurun_out/r2last_pytest.log
timeout 600 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2last_bench.log 2> gpurun_out/r2last_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2last_bench.log").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "spmv_ms", "non_spmv_ms_per_spmv")}, d["config"]["krylov_iterations_per_step"])
PY
