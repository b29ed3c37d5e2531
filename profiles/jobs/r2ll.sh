# round 2: ncu launch list of one outer-iteration window with the final library (per-program durations of the vector kernels)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 400 -c 140 --csv --log-file gpurun_out/launches_r2ll.csv python bench.py --ncu --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/r2ll_run.log 2>&1
tail -n 2 gpurun_out/r2ll_run.log | cut -c1-200
wc -l gpurun_out/launches_r2ll.csv
