# round 2, job u: lean stream loop variants of the SpMV against the production kernel
mkdir -p gpurun_out
MFB_SWEEP_VARIANTS=16,32,33,34,35,36,37,38,13 timeout 900 python bench.py --spmv-sweep > gpurun_out/r2u_sweep.log 2> gpurun_out/r2u_sweep.err
cut -c1-1500 gpurun_out/r2u_sweep.log
tail -n 3 gpurun_out/r2u_sweep.err
