python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fuzz.py tests/test_gpu_peer.py tests/test_gpu_sharded.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
python tools/select_probe.py resnet50 3 2>&1 | grep sampled | cut -c1-330
python tools/select_probe.py vit_l_16 3 2>&1 | grep sampled | cut -c1-330
python tools/select_trace.py resnet50 2>&1 | tail -1 | cut -c1-330
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],4),'kernel_ms',round(d['roofline']['kernel_ms'],4), {k: (round(v['us'],1), round(v['frac'],3)) for k,v in d['magnitude']['legs'].items()})"
