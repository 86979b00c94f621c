mkdir -p gpurun_out
IMPL=${1:-3}
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_lost_impl$IMPL.csv python tools/lost_probe.py 256 2 $IMPL > gpurun_out/ncu_lost_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lost_gram_tc2 -s 1 -c 1 -o gpurun_out/prof_lost_impl$IMPL -f python tools/lost_probe.py 256 2 $IMPL > gpurun_out/ncu_lost_full.log 2>&1; echo "ncu rc=$?"
