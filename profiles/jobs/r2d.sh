# round 2, job d (2 GPUs): multi-GPU parity over all cases and data planes, soak, N=2 bench line with dist_check;
# single-GPU tests of the new device mesh tables ride along
set -x
mkdir -p gpurun_out profiles/multigpu
nvidia-smi -L > gpurun_out/r2d_gpus.txt
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -v -s > gpurun_out/r2d_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_multigpu.log
timeout 300 python -m pytest tests/test_totalmesh_gpu.py -m gpu -q > gpurun_out/r2d_pytest_totalmesh.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_totalmesh.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2d_bench_n2.log 2> gpurun_out/r2d_bench_n2.err
MFB_KRYLOV_LEGACY=1 MFB_SPMV=row timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 2 --warmup 3 --no-extras --no-dist-check > gpurun_out/r2d_bench_n2_legacy.log 2> gpurun_out/r2d_bench_n2_legacy.err
grep -E "PASSED|FAILED|SKIPPED|passed|failed|DIST_OK|SOAK_OK|case=" gpurun_out/r2d_pytest_multigpu.log | tail -n 40
tail -n 3 gpurun_out/r2d_pytest_totalmesh.log
cut -c1-1500 gpurun_out/r2d_bench_n2.log | tail -n 2
tail -n 5 gpurun_out/r2d_bench_n2.err
