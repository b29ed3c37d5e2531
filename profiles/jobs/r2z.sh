# round 2, job z: two-rank tests (NH all planes, thermo, J2, soak) and the N=2 bench on the final library
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -k "neo_hookean or soak or j2_fused or thermo" > gpurun_out/r2z_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest_multigpu.log
tail -n 3 gpurun_out/r2z_pytest_multigpu.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2z_bench_n2.log 2> gpurun_out/r2z_bench_n2.err
tail -n 1 gpurun_out/r2z_bench_n2.log | cut -c1-400
