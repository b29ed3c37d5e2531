# round 2, job w: the 1-field (thermal conduction) workload in the bench
mkdir -p gpurun_out
timeout 600 python bench.py --workload thermal_conduction --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2w_thermal.log 2> gpurun_out/r2w_thermal.err
cut -c1-1200 gpurun_out/r2w_thermal.log | tail -n 1
tail -n 5 gpurun_out/r2w_thermal.err | cut -c1-300
