# round-1 final profile capture: every ncu command runs after the same command exited 0 without ncu
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-clocks"
$B > gpurun_out/plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_snip_score_sweep|k_snip_sample|k_select_finish|k_emit_masks" -s 8 -c 8 -o gpurun_out/prof_final_snip -f $B > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/lost_probe.py 256 5 3 > gpurun_out/plain2.log 2>&1 || { echo "lost probe failed"; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_lost_gram_tc2|k_lost_finish" -s 2 -c 2 -o gpurun_out/prof_final_lost -f python tools/lost_probe.py 256 2 3 > gpurun_out/ncu_full2.log 2>&1; echo "ncu lost rc=$?"
