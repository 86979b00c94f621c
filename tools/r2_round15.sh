mkdir -p gpurun_out
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | tail -4; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err | cut -c1-200
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2>&1 | tail -4
python tools/gram_error_probe.py 2>&1 | tail -4 > gpurun_out/r2_gram_error_probe.log; cat gpurun_out/r2_gram_error_probe.log | cut -c1-200
