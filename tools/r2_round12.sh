mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_gpu.log | cut -c1-300
python tools/select_probe.py resnet50 7 2>&1 | grep -E "sampled|world 1" | cut -c1-400
python tools/select_probe.py vit_l_16 3 2>&1 | grep -E "sampled|world 1" | cut -c1-400
python tools/lost_probe2.py 256 30 2>&1 | tail -2 | cut -c1-300
python __graft_entry__.py --smoke 2>&1 | tail -3
