mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 60 python tools/select_probe.py resnet50 6 2>&1 | tail -2
timeout 200 python bench.py --no-cpu-baseline --no-lost --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['result'])"
