"""globaltimer trace of the sample / sweep tails of one magnitude mask build (b200p_select_last_trace)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pruning_for_vision_representation_b200 import _lib as L
from pruning_for_vision_representation_b200.plan import ParamPlan
from pruning_for_vision_representation_b200.shapes import prunable_numels

model = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
dev = "cuda:0"
numels = prunable_numels(model); N = sum(numels)
plan = ParamPlan(numels, dev)
g = torch.Generator(device=dev); g.manual_seed(1)
w = [torch.randn(n, device=dev, generator=g) * 0.02 for n in numels]
plan.bind(L.SLOT_W, w)
mask = plan.new_mask()
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = L.load()
out = (ctypes.c_uint64 * 16)()
flush = (sys.argv[2] if len(sys.argv) > 2 else "flush") == "flush"
for it in range(4):
    if flush: junk.fill_(1)
    torch.cuda.synchronize()
    lib.b200p_select_last_trace(out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan.mask_build(L.KEY_ABS_W, N // 2, L.MODE_EXACT_K, mask); b.record(); torch.cuda.synchronize()
    lib.b200p_select_last_trace(out)
    t = [int(v) for v in out]
    t0 = t[0]
    rel = lambda i: (t[i] - t0) / 1e3 if t[i] else float("nan")
    print(f"{model} {'L2 flushed' if flush else 'no flush'} build {a.elapsed_time(b) * 1e3:.1f} us | sample tail: staged +{rel(1):.2f} bracket +{rel(2):.2f} done +{rel(3):.2f} | "
          f"sweep: first CTA start +{rel(8):.2f}, last flush issued +{rel(9):.2f}, tail entered +{rel(4):.2f}, staged +{rel(5):.2f}, window +{rel(6):.2f}, done +{rel(7):.2f} us",
          flush=True)
