# round 2, job q: lane-per-entry sweep kernel (component-major FP32 factors) against the stream kernel
mkdir -p gpurun_out
MFB_ILU_SWEEP=lane timeout 600 python -m pytest tests/test_krylov_gpu.py -m gpu -q -k ilu > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -n 3 gpurun_out/r2q_pytest.log
run() {
  env $1 timeout 600 python bench.py --ilu-only > gpurun_out/r2q_ilu_$2.log 2> gpurun_out/r2q_ilu_$2.err
  python - "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2q_ilu_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    r = d["ilu_only"][1]
    print(sys.argv[1], "solve_ms", round(r["solve_ms"], 1), "sweeps_ms", round(r["sweeps_ms_per_product"], 3), "its", r["krylov_iterations"], r["converged"], r["final_residual"], "fact+1", round(r["factorisation_plus_first_iteration_ms"], 1))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
  tail -n 2 gpurun_out/r2q_ilu_$2.err | cut -c1-200
}
run "MFB_ILU_SWEEP=lane MFB_ILU_RW=8" lane8
run "MFB_ILU_SWEEP=lane MFB_ILU_RW=4" lane4
run "MFB_ILU_SWEEP=lane MFB_ILU_RW=2" lane2
