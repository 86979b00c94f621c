python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "refresh or fused_equals" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n1_f.json 2> gpurun_out/r2_bench_n1_f.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n1_f.err | cut -c1-300
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_n1_f.json'))
print('value',d['value'],d['ms_per_step'],'frac',d['roofline']['frac'],'step_frac',d['roofline']['step_frac'],'kernel_ms',d['roofline']['kernel_ms'], 'launches', d['gpu_launches'])
print('contract',d['roofline']['contract']['value'],'fp32masks',d['with_fp32_masks']['ms_per_step'], 'e2e', d['e2e']['value'], d['result']['mask_sha256'][:16], d['roofline']['kernel_tails'])
P
