# round 2, job m: Pl_ILU timing split (factorisation, sweeps per product)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --ilu-only > gpurun_out/r2m_ilu.log 2> gpurun_out/r2m_ilu.err
cut -c1-2500 gpurun_out/r2m_ilu.log
tail -n 3 gpurun_out/r2m_ilu.err
