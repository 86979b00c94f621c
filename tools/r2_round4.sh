mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 ) > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest_gpu.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | tail -4; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
def short(v, depth=0):
    if isinstance(v, dict): return {k: short(x, depth+1) for k, x in v.items()} if depth < 2 else '{...}'
    if isinstance(v, list): return v[:4]
    if isinstance(v, str): return v[:60]
    return v
for k,v in d.items(): print(k, '=', short(v))
PY
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2>&1 | tail -4; cut -c1-1500 gpurun_out/r2_bench_ref.json
