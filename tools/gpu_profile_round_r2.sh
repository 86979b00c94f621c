# round-2 profile capture: every ncu command runs after the same command exited 0 without ncu
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-clocks --no-extra"
$B > gpurun_out/plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_snip_score_sweep|k_snip_sample|k_select_finish|k_fill_chunk_ptrs_multi" -s 10 -c 8 -o gpurun_out/prof_final_snip -f $B > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/lost_probe2.py 256 3 > gpurun_out/plain2.log 2>&1 || { echo "lost probe failed"; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_lost_gram_tc2|k_lost_finish" -s 8 -c 2 -o gpurun_out/prof_final_lost -f python tools/lost_probe2.py 256 2 > gpurun_out/ncu_full2.log 2>&1; echo "ncu lost rc=$?"
python tools/select_probe.py resnet50 2 > gpurun_out/plain3.log 2>&1 || { echo "select probe failed"; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_select_bracket|k_select_sample|k_select_finish|k_masked_sgd" -s 6 -c 6 -o gpurun_out/prof_final_select -f python tools/select_probe.py resnet50 2 > gpurun_out/ncu_full3.log 2>&1; echo "ncu select rc=$?"
ls -la gpurun_out/*.ncu-rep
