python -m pytest tests/test_gpu_multi.py tests/test_gpu_peer.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -1
N=${1:-2}
for lsu in 0 1; do
echo "== B200P_PUSH_LSU=$lsu"
B200P_PUSH_LSU=$lsu timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/comm_trace.py vit_l_16 2>&1 | grep "^rank" | tail -$N | sed "s/ (miss.*|| tail/ tail/" | head -3
done
