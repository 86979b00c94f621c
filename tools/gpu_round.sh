mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 60 python tools/select_probe.py resnet50 6 2>&1 | tail -2
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-clocks"
timeout 120 $CMD > gpurun_out/plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 200 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 120 $CMD > gpurun_out/plain2.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_score_multi|k_select_bracket|k_emit_masks|k_lost_gram_tc' -s 6 -c 6 -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
