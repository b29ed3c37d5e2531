# usage: bash profiles/ncu_job.sh <tag>   (run under gpurun; writes gpurun_out/*_<tag>.*)
# 1) launch list of one Newton step (cold-cache, serialised: compare SHARES), 2) --set full captures of the dominant kernels
TAG=${1:-r2}
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --ncu --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spmv_mr -s 20 -c 2 -o gpurun_out/prof_spmv_$TAG python bench.py --ncu --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_spmv_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mfb_b0_nl -c 1 -o gpurun_out/prof_elem_$TAG python bench.py --ncu --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_elem_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 30 -c 3 -o gpurun_out/prof_fused_$TAG python bench.py --ncu --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_fused_run.log 2>&1
ls -la gpurun_out | tail -n 8
