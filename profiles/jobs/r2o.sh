# round 2, job o: FP32-stored ILU factors -- tests, tuning variants, FP64 comparison
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_krylov_gpu.py -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -n 3 gpurun_out/r2o_pytest.log
run() {
  env $1 timeout 600 python bench.py --ilu-only > gpurun_out/r2o_ilu_$2.log 2> gpurun_out/r2o_ilu_$2.err
  python - "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2o_ilu_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    r = d["ilu_only"][1]
    print(sys.argv[1], "solve_ms", round(r["solve_ms"], 1), "sweeps_ms", round(r["sweeps_ms_per_product"], 3), "its", r["krylov_iterations"], r["converged"], r["final_residual"], "fact+1", round(r["factorisation_plus_first_iteration_ms"], 1))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
run MFB_ILU_CFG=0 f32_cfg0
run MFB_ILU_CFG=1 f32_cfg1
run MFB_ILU_CFG=2 f32_cfg2
run MFB_ILU_CFG=3 f32_cfg3
run MFB_ILU_CFG=4 f32_cfg4
run MFB_ILU_FP64=1 f64_cfg0
