mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2_pytest_gpu2.log 2>&1; tail -8 gpurun_out/r2_pytest_gpu2.log
python tools/select_probe.py resnet50 2>&1 | tail -3 | cut -c1-400
python tools/select_probe.py vit_l_16 2>&1 | tail -3 | cut -c1-400
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_n1_b.json'))
print('value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'step_frac',d['roofline']['step_frac'])
print('contract',d['roofline']['contract']['ms_per_step'], 'fp32masks', d['with_fp32_masks']['ms_per_step'])
for k,v in d['magnitude']['legs'].items(): print(k, v['us'], v['value'], v['frac'])
print('lost', d['lost']['value'], d['lost']['ms_per_step'], d['lost']['trace'])
P
