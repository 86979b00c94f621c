mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], d['e2e']['value'], d['lost']['value'])"; wc -l gpurun_out/bench_n2.json; tail -2 gpurun_out/bench_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --dist-mode sharded --no-lost > gpurun_out/bench_n2_sharded.json 2> gpurun_out/bench_n2s.err; echo "bench2s rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n2_sharded.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], d['e2e']['value'])"
