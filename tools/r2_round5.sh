mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest_gpu.log | cut -c1-300
python tools/select_probe.py resnet50 5 2>&1 | tail -2 | cut -c1-600
python tools/select_probe.py vit_l_16 3 2>&1 | tail -2 | cut -c1-600
python tools/lost_probe2.py 256 20 2>&1 | tail -2 | cut -c1-300
python tools/gram_error_probe.py 2>&1 | tail -4 | cut -c1-300
