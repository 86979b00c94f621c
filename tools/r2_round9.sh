mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 ) > gpurun_out/r2_pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest_sub.log | cut -c1-300
python tools/select_probe.py resnet50 7 2>&1 | tail -8 | cut -c1-600
python tools/lost_probe2.py 256 30 2>&1 | tail -2 | cut -c1-300
