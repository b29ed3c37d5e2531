# round 2, job h (2 GPUs): all two-rank tests again (J2 at the tighter comparison tolerance, fused idrs/bicgstab, deterministic
# scatter), N = 2 bench line with the rigorous dist_check
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -v -s > gpurun_out/r2h_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2h_bench_n2.log 2> gpurun_out/r2h_bench_n2.err
grep -E "PASSED|FAILED|passed|failed|^case=|SOAK_OK" gpurun_out/r2h_pytest_multigpu.log | cut -c1-330 | tail -n 40
cut -c1-300 gpurun_out/r2h_bench_n2.log | tail -n 1
tail -n 4 gpurun_out/r2h_bench_n2.err
