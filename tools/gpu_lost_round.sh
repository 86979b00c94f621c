mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lost.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 60 -k "tc2 or config3" > gpurun_out/pytest_lost.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_lost.log
for impl in 3 2; do for f in 0 4 1 3; do echo "impl=$impl dbg=$f"; B200P_LOST_DBG=$f timeout 60 python tools/lost_probe.py 256 20 $impl; done; done
