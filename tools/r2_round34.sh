python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/select_probe.py resnet50 3 2>&1 | grep sampled | cut -c1-330
python tools/select_probe.py vit_l_16 3 2>&1 | grep sampled | cut -c1-330
python tools/select_trace.py resnet50 2>&1 | tail -1 | cut -c1-330
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_n1_d.json'))
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],4),'kernel_ms',round(d['roofline']['kernel_ms'],4),'contract',round(d['roofline']['contract']['ms_per_step'],4))
print('  ', {k: (round(v['us'],1), round(v['value'],1), round(v['frac'],3)) for k,v in d['magnitude']['legs'].items()})
print('  lost', round(d['lost']['value']), d['lost']['ms_per_step'])
P
