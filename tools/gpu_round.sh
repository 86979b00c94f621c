mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lost.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 90 -k "lost" > gpurun_out/pytest_lost.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_lost.log
timeout 60 python tools/lost_probe.py 256 5
timeout 60 python tools/lost_probe.py 256 3 > gpurun_out/plain.log 2>&1 && timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_lost' -c 16 --csv --log-file gpurun_out/launches_lost.csv python tools/lost_probe.py 256 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"
