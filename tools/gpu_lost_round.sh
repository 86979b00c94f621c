mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lost.py -m gpu -q -x --timeout 60 -k "tc2d" 2>&1 | tail -2
timeout 60 python tools/lost_probe.py 256 20 3; timeout 60 python tools/lost_probe.py 256 20 3; timeout 60 python tools/lost_probe.py 1024 5 3
