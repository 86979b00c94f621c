"""Device-side timeline of the in-kernel collectives of one parameter-sharded magnitude build, under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/comm_trace.py [model]
Prints, per rank and collective: payload stores issued -> own fence done -> own flag stored -> last peer's flag seen -> leaving."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pruning_for_vision_representation_b200 import _lib as L
from pruning_for_vision_representation_b200.distributed import PeerShardedBuilder
from pruning_for_vision_representation_b200.plan import ParamPlan
from pruning_for_vision_representation_b200.shapes import prunable_numels

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
model = sys.argv[1] if len(sys.argv) > 1 else "vit_l_16"
numels = prunable_numels(model); n = sum(numels)
g = torch.Generator(device=dev).manual_seed(1)
w = torch.randn(n, device=dev, generator=g) * 0.02
plan = ParamPlan(numels, dev)
off = 0; views = []
for m in numels:
    views.append(w[off:off + m]); off += m
plan.bind(L.SLOT_W, views)
b = PeerShardedBuilder.from_process_group(plan)
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = (ctypes.c_uint64 * 64)()
names = ["hist", "gather", "mask", "barrier"]
for it in range(4):
    junk.fill_(1)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); b.magnitude_build(n // 2); e.record(); torch.cuda.synchronize()
    res = b.check()
    plan.lib.b200p_comm_trace(b.comm.handle, out)
    t = [int(v) for v in out]
    rows = []
    for ch in range(3):
        for par in range(2):
            s = t[(ch * 2 + par) * 8:(ch * 2 + par) * 8 + 8]
            if s[0] == 0: continue
            rows.append((s[0], f"{names[ch]}[{par}]", s))
    rows.sort()
    t0 = rows[0][0] if rows else 0
    msg = f"rank {rank} build {a.elapsed_time(e) * 1e3:.1f} us (miss {res['miss']}): "
    for _, name, s in rows:
        rel = lambda v: (v - t0) / 1e3 if v else float('nan')
        msg += f"| {name}: stores out +{rel(s[0]):.1f}, fence +{(s[1]-s[0])/1e3:.1f}, flag +{(s[2]-s[1])/1e3:.1f}, last peer seen +{(s[3]-s[2])/1e3:.1f}, leave +{(s[4]-s[3])/1e3:.1f}" + (f", AR entered {(s[5]-s[0])/1e3:.1f}, sums written +{(s[6]-s[4])/1e3:.1f}" if s[5] else "") + " "
    tl = t[(3 * 2) * 8:(3 * 2) * 8 + 8]
    if tl[0]:
        msg += "|| tail kernel: " + ", ".join(f"{nm} +{(tl[i + 1] - tl[i]) / 1e3:.1f}" for i, nm in enumerate(["window hist", "gather+key", "ties", "patch", "push", "final flag"]))
    if it >= 2:
        for r in range(world):
            if r == rank: print(msg, flush=True)
            dist.barrier()
dist.destroy_process_group()
