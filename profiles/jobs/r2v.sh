# round 2, job v: full GPU test-suite and the default bench on the final library
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2v_smoke.log 2>&1
timeout 1800 python bench.py > gpurun_out/r2v_bench.log 2> gpurun_out/r2v_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r2v_bench_ref.log 2> gpurun_out/r2v_bench_ref.err
tail -n 4 gpurun_out/r2v_pytest.log | cut -c1-300
tail -n 2 gpurun_out/r2v_smoke.log
cut -c1-300 gpurun_out/r2v_bench.log | tail -n 1
cut -c1-400 gpurun_out/r2v_bench_ref.log | tail -n 1
