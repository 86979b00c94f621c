set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python tools/microbench.py > gpurun_out/microbench_rn50.log 2>&1; grep -v "^{" gpurun_out/microbench_rn50.log | cut -c1-150
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
