# usage: bash profiles/ncu_elem.sh <tag>   -- one ncu --set full capture of the fused element kernel only
TAG=${1:-r1}
ncu --set full --clock-control none --import-source on -k regex:mfb_b0_nl -c 1 -o gpurun_out/prof_elem_$TAG python bench.py --ncu --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_elem_run.log 2>&1
