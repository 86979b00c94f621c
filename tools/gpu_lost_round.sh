mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_lost.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 120 -k "widths or config2 or tc2d" > gpurun_out/pytest_lost.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_lost.log
