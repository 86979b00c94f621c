"""LOST step timing, count-only (no A) vs A materialised, and how many boxes differ between the two on random keys."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pruning_for_vision_representation_b200 import object_discovery as OD
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
g = torch.Generator().manual_seed(0)
feats = torch.randn(B, 900, 384, generator=g).to(dev)
def timed(**kw):
    for _ in range(3):
        out = OD.lost_batched(feats, [30, 30], [16, 16], (3, 480, 480), **kw)
    torch.cuda.synchronize()
    import time
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); t0 = time.perf_counter()
    for _ in range(reps):
        out = OD.lost_batched(feats, [30, 30], [16, 16], (3, 480, 480), **kw)
    host = (time.perf_counter() - t0) / reps * 1e3
    b.record(); torch.cuda.synchronize()
    timed.host = host
    return a.elapsed_time(b) / reps, out
t0, o0 = timed(); h0 = timed.host
import ctypes
from pruning_for_vision_representation_b200 import _lib as L
tr = (ctypes.c_uint64 * 4)()
if L.load().b200p_lost_last_trace(tr) == 0:
    print(f"trace (us from Gram start): Gram end {(tr[1]-tr[0])/1e3:.1f}, first finish CTA ready {(tr[2]-tr[0])/1e3:.1f}, last finish end {(tr[3]-tr[0])/1e3:.1f}")
t1, o1 = timed(return_A=True)
same_deg = int((o0["degree"].flat.view(B, -1) == o1["degree"].flat.view(B, -1)).all(dim=1).sum())
same_box = int((o0["box"] == o1["box"]).all(dim=1).sum()); same_seed = int((o0["seed"] == o1["seed"]).sum())
print(f"B={B}: count-only {t0:.3f} ms ({B / t0 * 1e3:.0f} img/s, host enqueue {h0:.3f} ms) | with A {t1:.3f} ms ({B / t1 * 1e3:.0f} img/s) | "
      f"identical degrees {same_deg}/{B}, seeds {same_seed}/{B}, boxes {same_box}/{B}", flush=True)
