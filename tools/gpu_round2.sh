set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check rc=$?"; grep -v Warning gpurun_out/dist_check.log | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; cut -c1-2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; echo "bench1 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1b.json')); print(d['value'], d['lost'])"
