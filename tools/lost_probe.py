import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pruning_for_vision_representation_b200 import object_discovery as OD
if len(sys.argv) > 3:
    OD.DEFAULT_GRAM_IMPL = int(sys.argv[3])
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
feats = torch.randn(B, 900, 384, device=dev)
for _ in range(2):
    out = OD.lost_batched(feats, [30, 30], [16, 16], (3, 480, 480))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    out = OD.lost_batched(feats, [30, 30], [16, 16], (3, 480, 480))
b.record(); torch.cuda.synchronize()
print(f"B={B}: {a.elapsed_time(b)/reps:.3f} ms per batch, {B*reps/a.elapsed_time(b)*1e3:.0f} images/s")
