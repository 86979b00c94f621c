mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -q -x --timeout 120 ) > gpurun_out/r2_pytest_peer.log 2>&1; echo "pytest peer rc=$?"; tail -30 gpurun_out/r2_pytest_peer.log
