mkdir -p gpurun_out
( time python __graft_entry__.py --smoke ) 2>&1 | tail -6
( time timeout 600 python -m pytest tests -m gpu -q --timeout 120 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
( time python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | tail -4; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline']); print(d['clocks']); print(d['lost']['value'], d['lost']['e2e'], d['lost'].get('cpu_baseline'))"
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -4; cut -c1-400 gpurun_out/bench_ref.json
