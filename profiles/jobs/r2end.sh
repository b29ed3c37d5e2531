# round 2: last full GPU test-suite on the committed state (+ the light kernels switched off)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2end_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2end_pytest.log
tail -n 3 gpurun_out/r2end_pytest.log | cut -c1-300
MFB_NO_LIGHT=1 timeout 200 python -m pytest tests/test_krylov_gpu.py -m gpu -q > gpurun_out/r2end_pytest_nolight.log 2>&1; echo "nolight rc=$?"
tail -n 1 gpurun_out/r2end_pytest_nolight.log | cut -c1-200
