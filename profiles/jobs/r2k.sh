# round 2, job k: Pl_ILU elimination orders (hash vs colour classes) at 88^3
set -x
mkdir -p gpurun_out
MFB_ILU_ORDER=color timeout 600 python -m pytest tests/test_krylov_gpu.py -m gpu -q -k ilu > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
MFB_ILU_ORDER=color timeout 600 python bench.py --ilu-only > gpurun_out/r2k_ilu_color.log 2> gpurun_out/r2k_ilu_color.err
MFB_ILU_ORDER=hash timeout 600 python bench.py --ilu-only > gpurun_out/r2k_ilu_hash.log 2> gpurun_out/r2k_ilu_hash.err
tail -n 3 gpurun_out/r2k_pytest.log
cut -c1-1500 gpurun_out/r2k_ilu_color.log
cut -c1-1500 gpurun_out/r2k_ilu_hash.log
tail -n 3 gpurun_out/r2k_ilu_color.err
