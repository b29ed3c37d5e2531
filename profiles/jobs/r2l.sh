# round 2, job l: Pl_ILU with colour-class elimination order and stream-kernel sweeps at 88^3
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_krylov_gpu.py -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -n 3 gpurun_out/r2l_pytest.log
for cfg in "stream 8" "stream 16" "stream 4" "row 8"; do
  set -- $cfg
  MFB_ILU_SWEEP=$1 MFB_ILU_RW=$2 timeout 600 python bench.py --ilu-only > gpurun_out/r2l_ilu_$1_$2.log 2> gpurun_out/r2l_ilu_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2l_ilu_{sys.argv[1]}_{sys.argv[2]}.log").read().strip().splitlines()[-1])
    for r in d["ilu_only"]:
        print(sys.argv[1:], r["solve_ms"], r["krylov_iterations"], r["spmv_count"], r["dependency_levels"], r["converged"], r["final_residual"])
except Exception as e:
    print(sys.argv[1:], "unreadable", e)
PY
done
