# round 2, job n: sweep kernel tuning (unroll / occupancy / rows per warp)
mkdir -p gpurun_out
for cfg in 0 2 3 4; do
  MFB_ILU_CFG=$cfg timeout 600 python bench.py --ilu-only > gpurun_out/r2n_ilu_$cfg.log 2> gpurun_out/r2n_ilu_$cfg.err
  python - "$cfg" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2n_ilu_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    r = d["ilu_only"][1]
    print("cfg", sys.argv[1], "solve_ms", round(r["solve_ms"], 1), "sweeps_ms", round(r["sweeps_ms_per_product"], 3), "its", r["krylov_iterations"], r["converged"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
