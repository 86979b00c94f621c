mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lost.py tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -k "repeatable" > gpurun_out/pytest_rep.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_rep.log
