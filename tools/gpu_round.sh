mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac']); l=d['lost']; print(l['value'], l['ms_per_step'], l['roofline'], l['e2e'])"; tail -3 gpurun_out/bench_n1.err
