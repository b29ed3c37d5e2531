# round 2, job p: where the time of the ILU sweeps goes -- level sizes, per-launch durations, one full capture
mkdir -p gpurun_out
MFB_ILU_VERBOSE=1 MFB_ILU_CFG=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_sweep_mr -s 126 -c 63 --csv --log-file gpurun_out/r2p_sweep_launches.csv python bench.py --ilu-only > gpurun_out/r2p_run1.log 2> gpurun_out/r2p_run1.err
grep "mfb ilu" gpurun_out/r2p_run1.err | head -2
MFB_ILU_CFG=1 ncu --set full --clock-control none --import-source on -k regex:k_sweep_mr -s 128 -c 3 -o gpurun_out/prof_sweep_r2p python bench.py --ilu-only > gpurun_out/r2p_run2.log 2>&1
ls -la gpurun_out | grep r2p
