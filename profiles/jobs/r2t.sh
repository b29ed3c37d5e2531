# round 2, job t: pre-scaled upper rows (no block product in the sweep epilogue); Krylov tests on every ILU path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_krylov_gpu.py -m gpu -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -n 2 gpurun_out/r2t_pytest.log
for e in MFB_ILU_FP64=1 MFB_ILU_SWEEP=row MFB_ILU_UNPACKED=1 MFB_ILU_ORDER=hash; do
  env $e timeout 600 python -m pytest tests/test_krylov_gpu.py -m gpu -q -k ilu > gpurun_out/r2t_pytest_$e.log 2>&1; echo "$e pytest rc=$?"
done
run() {
  env $1 timeout 600 python bench.py --ilu-only > gpurun_out/r2t_ilu_$2.log 2> gpurun_out/r2t_ilu_$2.err
  python - "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2t_ilu_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    r = d["ilu_only"][1]
    print(sys.argv[1], "solve_ms", round(r["solve_ms"], 1), "sweeps_ms", round(r["sweeps_ms_per_product"], 3), "spmv_ms", round(r["spmv_ms"], 3), "its", r["krylov_iterations"], r["converged"], r["final_residual"], "fact+1", round(r["factorisation_plus_first_iteration_ms"], 1))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
run "MFB_ILU_CFG=0" cfg0
run "MFB_ILU_CFG=1" cfg1
run "MFB_ILU_CFG=4" cfg4
