"""Short driver for ncu: a few mask builds (SNIP select+emit, magnitude select+emit) on the ResNet-50 set."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pruning_for_vision_representation_b200 import _lib as L
from pruning_for_vision_representation_b200.plan import ParamPlan
from pruning_for_vision_representation_b200.shapes import prunable_numels

model = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = "cuda:0"
numels = prunable_numels(model); N = sum(numels)
plan = ParamPlan(numels, dev)
g = torch.Generator(device=dev); g.manual_seed(1)
mk = lambda scale: [torch.randn(n, device=dev, generator=g) * scale for n in numels]
w, gr, sc = mk(0.02), mk(1e-3), mk(1.0)
plan.bind(L.SLOT_W, w).bind(L.SLOT_G, gr).bind(L.SLOT_SCORE, sc)
plan.score_accumulate(False)
mask = plan.new_mask(); old = plan.new_mask()
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn):
    ts = []
    for _ in range(reps):
        junk.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]
for impl in ("sampled", "exact"):
    plan.set_select_impl(impl)
    t1 = timed(lambda: plan.select_kth(L.KEY_SCORE, int(N * 0.9), L.MODE_SNIP_STRICT)); r1 = plan.result()
    t2 = timed(lambda: plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, mask))
    t3 = timed(lambda: plan.select_kth(L.KEY_ABS_W, N // 2, L.MODE_EXACT_K)); r3 = plan.result()
    t4 = timed(lambda: plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, mask))
    old.copy_(mask)
    t5 = timed(lambda: plan.select_kth(L.KEY_ABS_W, N // 10, L.MODE_EXACT_K, old)); r5 = plan.result()
    def full():
        plan.mask_build(L.KEY_ABS_W, N // 2, L.MODE_EXACT_K, mask)
    t6 = timed(full)
    t7 = timed(lambda: plan.mask_build(L.KEY_SCORE, int(N * 0.9), L.MODE_SNIP_STRICT, mask))
    print(f"{model} {impl}: select_score {t1:.1f} us (passes {r1['passes_full']}, cand {r1['collected']}) emit_snip {t2:.1f} | "
          f"select_absw {t3:.1f} (passes {r3['passes_full']}) emit_mag {t4:.1f} | round2 select {t5:.1f} (passes {r5['passes_full']}) | "
          f"magnitude build {t6:.1f} us = {N / t6 / 1e3:.1f} Gparams/s | snip select+emit {t7:.1f} us", flush=True)

# the sharded sequence on ONE rank (world = 1: every in-kernel collective degenerates to a local copy): its fixed cost against
# mask_build on the same data, stage by stage
from pruning_for_vision_representation_b200.distributed import PeerComm, PeerShardedBuilder
plan.set_select_impl("sampled")
comm = PeerComm(plan, 0, 1)
b = PeerShardedBuilder(plan, comm)
tb = timed(lambda: b.magnitude_build(N // 2))
print(f"{model} sharded sequence, world 1: magnitude build {tb:.1f} us ({b.check()})", flush=True)
names = ["sample", "sweep", "finish", "ties", "emit", "push"]
for bit, name in enumerate(names):
    # run the earlier stages untimed, then time this one
    def upto():
        for j in range(bit):
            b._build(L.KEY_ABS_W, None, N // 2, L.MODE_EXACT_K, stages=1 << j)
    ts = []
    for _ in range(reps):
        junk.fill_(1); upto()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(); b._build(L.KEY_ABS_W, None, N // 2, L.MODE_EXACT_K, stages=1 << bit); b_.record(); torch.cuda.synchronize()
        ts.append(a_.elapsed_time(b_) * 1e3)
        for j in range(bit + 1, 6):
            b._build(L.KEY_ABS_W, None, N // 2, L.MODE_EXACT_K, stages=1 << j)
        torch.cuda.synchronize()
    print(f"   stage {name}: {sorted(ts)[len(ts) // 2]:.1f} us", flush=True)
