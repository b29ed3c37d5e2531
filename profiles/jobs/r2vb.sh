# round 2: virtual-block light kernels (same sums bit for bit, one physical wave): Krylov tests + one timed step
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_krylov_gpu.py -m gpu -q > gpurun_out/r2vb_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2vb_pytest.log
tail -n 2 gpurun_out/r2vb_pytest.log | cut -c1-200
timeout 300 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2vb_bench.log 2> gpurun_out/r2vb_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2vb_bench.log").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "spmv_ms", "non_spmv_ms_per_spmv")}, d["config"]["krylov_iterations_per_step"], d["config"]["final_residual_max"])
PY
