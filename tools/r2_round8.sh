mkdir -p gpurun_out
python tools/select_probe.py resnet50 7 2>&1 | tail -9 | cut -c1-400
python tools/select_probe.py resnet152 5 2>&1 | tail -8 | cut -c1-400
python tools/lost_probe2.py 256 30 2>&1 | tail -2 | cut -c1-300
( timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_training.py tests/test_gpu_lost.py -m gpu -q --timeout 300 ) > gpurun_out/r2_pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest_sub.log | cut -c1-300
