mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -q --timeout 90 > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_train.log
