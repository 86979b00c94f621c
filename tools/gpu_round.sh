mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 200 > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_full.log
