mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for i in 1 2 3; do timeout 60 python tools/select_probe.py resnet50 6 > gpurun_out/select_probe.log 2>&1 ; echo "probe rc=$?"; cat gpurun_out/select_probe.log; done
timeout 60 python tools/select_probe.py resnet50 2 > gpurun_out/plain.log 2>&1 && timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_select|k_emit|k_tie' -c 30 --csv --log-file gpurun_out/launches_select.csv python tools/select_probe.py resnet50 2 > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"
