# round 2: the ILU application against its oracle restatement
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_krylov_gpu.py -m gpu -q -k "ilu" > gpurun_out/r2ilu_oracle.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ilu_oracle.log
tail -n 25 gpurun_out/r2ilu_oracle.log | cut -c1-300
