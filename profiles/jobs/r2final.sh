# round 2, final check of the library as committed: full GPU test-suite, smoke, default bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2final_smoke.log 2>&1
timeout 1800 python bench.py > gpurun_out/r2final_bench.log 2> gpurun_out/r2final_bench.err
tail -n 4 gpurun_out/r2final_pytest.log | cut -c1-300
tail -n 1 gpurun_out/r2final_smoke.log
cut -c1-300 gpurun_out/r2final_bench.log | tail -n 1
