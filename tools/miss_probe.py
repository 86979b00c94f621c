"""Which selects fall back to the exact path (bracket miss)?  Sweeps on synthetic and default-init weights, iterative rounds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.utils.prune as prune
from pruning_for_vision_representation_b200 import _lib as L
from pruning_for_vision_representation_b200.plan import ParamPlan
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda:0")
def views(flat, numels):
    out, off = [], 0
    for n in numels:
        out.append(flat[off:off + n]); off += n
    return out
# 1. small synthetic set, plain vs reuse
numels = [4096 * 90 + 3, 4096 * 40, 12345]; n = sum(numels)
g = torch.Generator(device=dev).manual_seed(9)
wt = torch.randn(n, device=dev, generator=g) * 0.02
for reuse in (False, True):
    p = ParamPlan(numels, dev)
    if reuse: p.reuse_sample()
    p.bind(L.SLOT_W, views(wt, numels))
    for s in (0.5, 0.9, 0.2, 0.99):
        m = p.new_mask(); p.mask_build(L.KEY_ABS_W, round(s * n), L.MODE_EXACT_K, m); r = p.result()
        print(f"small reuse={reuse} s={s}: miss={r['miss']} passes={r['passes_full']} collected={r['collected']} n_equal={r['n_equal']}", flush=True)
# 2. ViT-B iterative rounds on default init
w, numels, src = bench.model_weights_cpu("vit_b_16"); n = sum(numels)
wd = w.to(dev); p = ParamPlan(numels, dev); p.bind(L.SLOT_W, views(wd, numels))
old, n_alive = None, n
for rnd in range(14):
    kk = prune._compute_nparams_toprune(0.2, n_alive)
    new = p.new_mask()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); p.mask_build(L.KEY_ABS_W, kk, L.MODE_EXACT_K, new, old); b.record(); torch.cuda.synchronize()
    r = p.result()
    print(f"vit_b round {rnd}: {a.elapsed_time(b)*1e3:.0f} us miss={r['miss']} passes={r['passes_full']} collected={r['collected']} n_equal={r['n_equal']} quota={r['quota']} n_alive={n_alive}", flush=True)
    old, n_alive = new, n_alive - kk
