# round 2, job r: stream sweeps with FP32 values held in float registers, deeper unroll variants
mkdir -p gpurun_out
run() {
  env $1 timeout 600 python bench.py --ilu-only > gpurun_out/r2r_ilu_$2.log 2> gpurun_out/r2r_ilu_$2.err
  python - "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2r_ilu_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    r = d["ilu_only"][1]
    print(sys.argv[1], "solve_ms", round(r["solve_ms"], 1), "sweeps_ms", round(r["sweeps_ms_per_product"], 3), "its", r["krylov_iterations"], r["converged"], r["final_residual"], "fact+1", round(r["factorisation_plus_first_iteration_ms"], 1))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
  tail -n 2 gpurun_out/r2r_ilu_$2.err | cut -c1-200
}
for c in 0 2 3 4 5 6; do run "MFB_ILU_CFG=$c" cfg$c; done
