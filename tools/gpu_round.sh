mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lost.py -m gpu -q -x --timeout 60 > gpurun_out/pytest_lost.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_lost.log
timeout 60 python tools/lost_probe.py 256 5; timeout 60 python tools/lost_probe.py 1024 3
timeout 60 python tools/lost_probe.py 256 3 > gpurun_out/plain.log 2>&1 && timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_lost' -c 14 --csv --log-file gpurun_out/launches_lost.csv python tools/lost_probe.py 256 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"
