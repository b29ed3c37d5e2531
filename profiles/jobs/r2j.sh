# round 2, job j: 8-GPU strong-scaling point of the bench (what the driver's SCALE run does at N=8), with the dist_check and the breakdown
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2j_bench_n8.log 2> gpurun_out/r2j_bench_n8.err
echo "rc=$?"
tail -n 1 gpurun_out/r2j_bench_n8.log | cut -c1-1500
tail -n 5 gpurun_out/r2j_bench_n8.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2j_bench_n4.log 2> gpurun_out/r2j_bench_n4.err
echo "rc=$?"
tail -n 1 gpurun_out/r2j_bench_n4.log | cut -c1-600
