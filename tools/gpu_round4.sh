mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check4.log 2>&1; echo "dist_check rc=$?"; grep -v "Warning\|^\*\|OMP" gpurun_out/dist_check4.log | tail -9
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench4 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n4.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['lost']['value'])"; tail -3 gpurun_out/bench_n4.err
