"""Phase trace of the LOST finish kernel for a few images of one count-only batch (b200p_lost_finish_trace)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pruning_for_vision_representation_b200 import object_discovery as OD
from pruning_for_vision_representation_b200 import _lib as L
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(0)
feats = torch.randn(B, 900, 384, generator=g).to(dev)
lib = L.load()
out = (ctypes.c_uint64 * 16)()
names = ["deg+hist+seed", "cut-off", "potentials", "A[seed,p]+sim", "order", "vsum", "M", "component+box"]
run = lambda: OD.lost_batched(feats, [30, 30], [16, 16], (3, 480, 480))
for _ in range(3): run()
tr4 = (ctypes.c_uint64 * 4)()
for img in (0, 100, 200, 255):
    lib.b200p_lost_finish_trace(img, out)       # select the image for the next call
    run(); torch.cuda.synchronize()
    lib.b200p_lost_last_trace(tr4)
    lib.b200p_lost_finish_trace(img, out)
    t = [int(v) for v in out]
    d = [(t[i + 1] - t[i]) / 1e3 for i in range(8)]
    print(f"image {img}: ready at {(t[0] - tr4[0]) / 1e3:.1f} us after Gram start (Gram end {(tr4[1] - tr4[0]) / 1e3:.1f}), total {(t[8] - t[0]) / 1e3:.1f} us | " +
          " | ".join(f"{n} {v:.1f}" for n, v in zip(names, d)), flush=True)
