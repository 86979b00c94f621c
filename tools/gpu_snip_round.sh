mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for mode in sweep fused; do
timeout 300 python bench.py --no-lost --no-e2e --no-cpu-baseline --score-mode $mode > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; echo "bench $mode rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_$mode.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['step_algorithmic_GBps'], d['result'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_sweep.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-clocks --no-lost > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"
