# round 2, job s: does the swept vector stay in L2? evict_last cache hints / persisting window
mkdir -p gpurun_out
run() {
  env $1 timeout 600 python bench.py --ilu-only > gpurun_out/r2s_ilu_$2.log 2> gpurun_out/r2s_ilu_$2.err
  python - "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2s_ilu_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    r = d["ilu_only"][1]
    print(sys.argv[1], "solve_ms", round(r["solve_ms"], 1), "sweeps_ms", round(r["sweeps_ms_per_product"], 3), "spmv_ms", round(r["spmv_ms"], 3), "its", r["krylov_iterations"], r["converged"], r["final_residual"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
  grep "mfb ilu" gpurun_out/r2s_ilu_$2.err | head -2
}
run "MFB_ILU_CFG=2" hint_minb4
run "MFB_ILU_CFG=3" hint_minb3
run "MFB_ILU_CFG=6 MFB_ILU_L2WIN=1" window
run "MFB_ILU_CFG=6" base
