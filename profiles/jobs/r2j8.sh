# round 2: 8-GPU strong-scaling point of the final library
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2j8_bench_n8.log 2> gpurun_out/r2j8_bench_n8.err
echo "rc=$?"
tail -n 1 gpurun_out/r2j8_bench_n8.log | cut -c1-300
