mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 60 python tools/select_probe.py resnet50 6 > gpurun_out/select_probe.log 2>&1 ; echo "probe rc=$?"; cat gpurun_out/select_probe.log
timeout 60 python tools/select_probe.py resnet50 2 > gpurun_out/plain.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_select_bracket|k_emit_masks' -s 1 -c 2 -o gpurun_out/prof_select3 python tools/select_probe.py resnet50 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
