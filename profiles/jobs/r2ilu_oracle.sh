# round 2: the ILU application against its oracle restatement, on every A/B path of the library
mkdir -p gpurun_out
for e in MFB_ILU_FP64=1 MFB_ILU_ORDER=hash MFB_ILU_UNPACKED=1 "MFB_ILU_SWEEP=row MFB_ILU_FP64=1"; do
  env $e timeout 120 python -m pytest tests/test_krylov_gpu.py -m gpu -q -k "oracle_restatement" > "gpurun_out/r2ilu_oracle_$(echo $e | tr ' =' '__').log" 2>&1
  echo "$e rc=$? $(tail -n 1 "gpurun_out/r2ilu_oracle_$(echo $e | tr ' =' '__').log" | cut -c1-200)"
done
