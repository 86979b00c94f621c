"""Max |A - A64| / (|k_i| |k_j|) of every Gram implementation for a few key widths (the parity bar is 1e-5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pruning_for_vision_representation_b200 import object_discovery as OD, _lib as L
dev = torch.device("cuda:0")
for d in (384, 768, 1024, 2048):
    g = torch.Generator().manual_seed(d)
    f = torch.randn(2, 400, d, generator=g)
    f64 = f[0].double().numpy(); A64 = f64 @ f64.T; nrm = np.sqrt(np.diag(A64)); sc = np.outer(nrm, nrm)
    row = []
    for name, impl in (("ffma", L.LOST_GRAM_FFMA), ("tc", L.LOST_GRAM_TC), ("tc2", L.LOST_GRAM_TC2), ("tc2d", L.LOST_GRAM_TC2D)):
        out = OD.lost_batched(f.to(dev), [20, 20], [16, 16], (3, 320, 320), return_A=True, gram_impl=impl)
        A = out["A"][0].cpu().numpy()
        row.append(f"{name} {np.max(np.abs(A - A64) / sc):.2e}")
    a32 = (f[0].to(dev) @ f[0].to(dev).T).cpu().numpy()
    row.append(f"torch fp32 matmul {np.max(np.abs(a32 - A64) / sc):.2e}")
    print(f"d={d}: " + " | ".join(row), flush=True)
