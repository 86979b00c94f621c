mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -2 ) 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('value',d['value'],d['ms_per_step'],'frac',d['roofline']['frac'],'step_frac',d['roofline']['step_frac'],'kernel_ms',d['roofline']['kernel_ms'])
print('contract',d['roofline']['contract']['value'],d['roofline']['contract']['frac'],'fp32masks',d['with_fp32_masks']['ms_per_step'],d['with_fp32_masks']['value'])
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['same_mask_as_gpu'])
for k,v in d['magnitude']['legs'].items(): print(k, round(v['us'],1), round(v['value'],1), round(v['frac'],3))
l=d['lost']; print('lost',l['value'],l['ms_per_step'],l['roofline']['kernel_us'],l['voc_mix']['value'],l['voc_mix']['parity']['ok'])
P
