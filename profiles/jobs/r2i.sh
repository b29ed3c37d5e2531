# round 2, job i: packed ILU sweeps -- Krylov tests, then the bench's ILU extra (packed vs first version)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_krylov_gpu.py -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
timeout 900 python bench.py --steps 1 --warmup 3 > gpurun_out/r2i_bench.log 2> gpurun_out/r2i_bench.err
tail -n 3 gpurun_out/r2i_pytest.log
python - <<'PY'
import json
for f in ("gpurun_out/r2i_bench.log",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], json.dumps(d.get("extra", {}).get("neo_hookean_pl_ilu"))[:900])
    except Exception as e:
        print(f, "unreadable", e)
PY
