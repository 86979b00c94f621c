"""Generate the golden fixtures by running the UNMODIFIED reference in the build container.

    PYTHONDONTWRITEBYTECODE=1 WANDB_MODE=disabled python tests/golden/make_golden.py

Imports `train` and `object_discovery` from /root/reference (read-only) and stores the
inputs/outputs of the hot-path functions as small .npz/.json files next to this script.
/root/reference does not exist on the GPU box, so tests only ever read the fixtures.
`skimage` (imported by the reference's datasets.py, not installed here) is stubbed; nothing on
the hot path uses it.  torch.argsort is pinned to stable=True while `lost()` runs (tie policy,
SURVEY §8c); the un-pinned outputs are stored beside it for information.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B200P_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
os.environ.setdefault("WANDB_MODE", "disabled")
sys.path.insert(0, REF)
for name in ("skimage", "skimage.io"):
    sys.modules.setdefault(name, types.ModuleType(name))

import train as ref_train                      # noqa: E402  (/root/reference/train.py)
import object_discovery as ref_od              # noqa: E402  (/root/reference/object_discovery.py)
import torch.nn.utils.prune as prune           # noqa: E402


def prunable(model):
    return [(n, m) for n, m in model.named_modules() if isinstance(m, (nn.Conv2d, nn.Linear))]


class TinyNet(nn.Module):
    """conv / conv / linear / linear with awkward sizes (partial chunks, numel % 4 != 0)."""

    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv2d(3, 13, 3, padding=1)        # 351
        self.c2 = nn.Conv2d(13, 24, 3, padding=1)       # 2808
        self.pool = nn.AdaptiveAvgPool2d(4)
        self.f1 = nn.Linear(24 * 16, 67)                # 25728  (> 1 chunk of 4096)
        self.f2 = nn.Linear(67, 10)                     # 670

    def forward(self, x):
        x = torch.relu(self.c1(x))
        x = torch.relu(self.c2(x))
        x = self.pool(x).flatten(1)
        return self.f2(torch.relu(self.f1(x)))


def masks_of(model):
    return [m.weight_mask.detach().numpy().astype(np.uint8) for _, m in prunable(model)]


def origs_of(model):
    return [(m.weight_orig if hasattr(m, "weight_orig") else m.weight).detach().numpy().copy() for _, m in prunable(model)]


def sha(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.packbits(np.asarray(a).reshape(-1).astype(bool), bitorder="little").tobytes())
    return h.hexdigest()


def gen_magnitude_tiny():
    torch.manual_seed(7)
    model = TinyNet()
    # plant ties at the threshold region: quantise one layer so many |w| coincide
    with torch.no_grad():
        model.f1.weight.copy_((model.f1.weight * 64).round() / 64)
    out = {f"w{i}": w for i, w in enumerate(origs_of(model))}
    ref_train.magnitude_pruning(model, 0.5)
    for i, m in enumerate(masks_of(model)):
        out[f"m1_{i}"] = m
    s1 = ref_train.compute_sparsity_global(model)
    ref_train.magnitude_pruning(model, 0.2)
    for i, m in enumerate(masks_of(model)):
        out[f"m2_{i}"] = m
    s2 = ref_train.compute_sparsity_global(model)
    out["sparsity"] = np.array([s1, s2], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "magnitude_tiny.npz"), **out)
    print("magnitude_tiny: sparsity", s1, s2)


def gen_magnitude_tiefree():
    """continuous weights: the tied set has size 1, masks must match bit for bit."""
    torch.manual_seed(11)
    model = TinyNet()
    out = {f"w{i}": w for i, w in enumerate(origs_of(model))}
    sp = []
    for r, amount in enumerate((0.3, 0.2, 0.2, 0.5)):
        ref_train.magnitude_pruning(model, amount)
        for i, m in enumerate(masks_of(model)):
            out[f"m{r}_{i}"] = m
        sp.append(ref_train.compute_sparsity_global(model))
    out["amounts"] = np.array([0.3, 0.2, 0.2, 0.5])
    out["sparsity"] = np.array(sp)
    np.savez_compressed(os.path.join(HERE, "magnitude_tiefree.npz"), **out)
    print("magnitude_tiefree: sparsity", sp)


def gen_snip_tiny():
    torch.manual_seed(3)
    model = TinyNet()
    x = torch.randn(8, 3, 16, 16)
    y = torch.randint(0, 10, (8,))
    crit = nn.CrossEntropyLoss()
    out = {f"w{i}": w for i, w in enumerate(origs_of(model))}
    # gradients exactly as the reference's hooks see them (train.py:258-275)
    model.zero_grad()
    crit(model(x), y).backward()
    for i, (_, m) in enumerate(prunable(model)):
        out[f"g{i}"] = m.weight.grad.detach().numpy().copy()
    model.zero_grad()
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref_train.snip_pruning(model, [(x, y)], torch.device("cpu"), crit, target_sparsity=0.9)
    thr_line = [l for l in buf.getvalue().splitlines() if l.startswith("SNIP threshold:")][0]
    out["threshold"] = np.array([float(thr_line.split(":")[1])], dtype=np.float64)
    for i, m in enumerate(masks_of(model)):
        out[f"m{i}"] = m
    out["sparsity"] = np.array([ref_train.compute_sparsity_global(model)])
    out["x"] = x.numpy(); out["y"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "snip_tiny.npz"), **out)
    print("snip_tiny:", thr_line, "sparsity", out["sparsity"])


def gen_resnet18_known_answer():
    import torchvision
    torch.manual_seed(1)
    model = torchvision.models.get_model("resnet18", weights=None, num_classes=1000)
    n = sum(m.weight.numel() for _, m in prunable(model))
    flat = torch.cat([m.weight.detach().abs().view(-1) for _, m in prunable(model)])
    k = round(0.5 * n)
    kth = torch.sort(flat)[0][k - 1].item()
    n_equal = int((flat == kth).sum())
    tied1 = torch.nonzero(flat == kth).view(-1).tolist()
    ref_train.magnitude_pruning(model, 0.5)
    s1 = ref_train.compute_sparsity_global(model)
    m1 = masks_of(model)
    kept1 = [int(m.sum()) for m in m1]
    flat_m1 = np.concatenate([m.reshape(-1) for m in m1])
    # second round: 20 % of the survivors
    eff = torch.cat([m.weight.detach().abs().view(-1) for _, m in prunable(model)])
    alive = torch.from_numpy(flat_m1.astype(bool))
    k2 = round(0.2 * int(alive.sum()))
    kth2 = torch.sort(eff[alive])[0][k2 - 1].item()
    tied2 = torch.nonzero((eff == kth2) & alive).view(-1).tolist()
    ref_train.magnitude_pruning(model, 0.2)
    s2 = ref_train.compute_sparsity_global(model)
    m2 = masks_of(model)
    flat_m2 = np.concatenate([m.reshape(-1) for m in m2])
    info = {"model": "resnet18", "seed": 1, "N": n, "k": k, "kth_abs": kth, "n_equal_at_kth": n_equal,
            "sparsity_after_0.5": s1, "sparsity_after_0.5_then_0.2": s2,
            "kept_per_tensor_round1": kept1, "kept_per_tensor_round2": [int(m.sum()) for m in m2],
            "mask_sha256_round1": sha(m1), "mask_sha256_round2": sha(m2),
            "n_less_round1": int((flat < kth).sum()),
            "tied_flat_index_round1": tied1, "tied_ref_mask_round1": [int(flat_m1[i]) for i in tied1],
            "k_round2": k2, "kth_abs_round2": kth2,
            "tied_flat_index_round2": tied2, "tied_ref_mask_round2": [int(flat_m2[i]) for i in tied2],
            "torch": torch.__version__, "torchvision": torchvision.__version__}
    with open(os.path.join(HERE, "resnet18_magnitude.json"), "w") as f:
        json.dump(info, f, indent=1)
    print("resnet18:", {k_: info[k_] for k_ in ("N", "k", "kth_abs", "n_equal_at_kth", "sparsity_after_0.5")})


def run_lost(feats, dims, scales, size, k_patches, stable):
    orig = torch.argsort
    if stable:
        torch.argsort = lambda t, **kw: orig(t, stable=True, **{k: v for k, v in kw.items() if k != "stable"})
    try:
        pred, A, scores, seed = ref_od.lost(feats, dims, scales, size, k_patches=k_patches)
    finally:
        torch.argsort = orig
    return np.asarray(pred), A.numpy(), scores.numpy(), int(seed)


def gen_lost():
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.lost_oracle import planted_object_feats
    cases = {}
    # random Gaussian keys, config 3 shape
    torch.manual_seed(0)
    f = torch.randn(1, 900, 384)
    cases["random900"] = (f.numpy()[0], [30, 30], [16, 16], (3, 480, 480), 100)
    # planted objects, a few shapes (VOC-like non-square grids)
    rng = np.random.default_rng(5)
    cases["planted900"] = (planted_object_feats(rng), [30, 30], [16, 16], (3, 480, 480), 100)
    cases["planted768"] = (planted_object_feats(rng, grid=(24, 32), rows=(5, 15), cols=(12, 30)), [24, 32], [16, 16],
                           (3, 375, 500), 100)
    cases["planted_small"] = (planted_object_feats(rng, grid=(7, 9), d=64, rows=(2, 5), cols=(3, 8)), [7, 9], [16, 16],
                              (3, 112, 144), 20)
    out = {}
    meta = {}
    for name, (feats, dims, scales, size, kp) in cases.items():
        ft = torch.from_numpy(feats)[None]
        pred_s, A, scores, seed_s = run_lost(ft, dims, scales, size, kp, stable=True)
        pred_u, _, _, seed_u = run_lost(ft, dims, scales, size, kp, stable=False)
        out[f"{name}_feats"] = feats.astype(np.float32)
        out[f"{name}_degree"] = (-scores).astype(np.int32)
        out[f"{name}_pred"] = pred_s.astype(np.int64)
        out[f"{name}_A_probe"] = A[:4, :6].astype(np.float32)
        meta[name] = {"dims": dims, "scales": scales, "init_image_size": list(size), "k_patches": kp,
                      "seed": seed_s, "pred": [int(v) for v in pred_s],
                      "unpinned_seed": seed_u, "unpinned_pred": [int(v) for v in pred_u]}
        print(name, meta[name])
    np.savez_compressed(os.path.join(HERE, "lost_cases.npz"), **out)
    with open(os.path.join(HERE, "lost_cases.json"), "w") as fjs:
        json.dump(meta, fjs, indent=1)


def gen_vit_producer():
    """Pins of the LOST feature producer (SURVEY f-4): (a) the reference's own position-embedding interpolation
    (vision_transformer.py:781-858) on a 14x14 -> 13x17 and -> 30x30 grid; (b) the k extraction of
    main_lost_original.py:251-263, run verbatim on a random qkv tensor."""
    from collections import OrderedDict
    import vision_transformer as ref_vit          # /root/reference/vision_transformer.py
    torch.manual_seed(11)
    pe = torch.randn(1, 1 + 14 * 14, 12)
    out = {"pos_embed": pe.numpy()}
    for (h, w) in ((13, 17), (30, 30)):
        st = OrderedDict({"encoder.pos_embedding": pe.clone()})
        new = ref_vit.interpolate_embeddings((h * 16, w * 16), 16, st)["encoder.pos_embedding"]
        out[f"interp_{h}x{w}"] = new.numpy()
    nb_im, nb_tokens, nh, D = 2, 1 + 35, 6, 48
    feat_out = {"qkv": torch.randn(nb_im, nb_tokens, 3 * D)}
    qkv = (feat_out["qkv"].reshape(nb_im, nb_tokens, 3, nh, -1 // nh).permute(2, 0, 3, 1, 4))      # main_lost_original.py:251-255, verbatim
    q, k, v = qkv[0], qkv[1], qkv[2]
    k = k.transpose(1, 2).reshape(nb_im, nb_tokens, -1)                                             # :257
    out["qkv"] = feat_out["qkv"].numpy()
    out["k_feats"] = k[:, 1:, :].contiguous().numpy()                                               # :263
    np.savez_compressed(os.path.join(HERE, "vit_producer.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["magnitude_tiny", "magnitude_tiefree", "snip_tiny", "resnet18", "lost", "vit"]
    if "vit" in which: gen_vit_producer()
    if "magnitude_tiny" in which: gen_magnitude_tiny()
    if "magnitude_tiefree" in which: gen_magnitude_tiefree()
    if "snip_tiny" in which: gen_snip_tiny()
    if "resnet18" in which: gen_resnet18_known_answer()
    if "lost" in which: gen_lost()
