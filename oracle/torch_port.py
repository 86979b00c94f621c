"""CPU timing port of the reference's mask build — TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python over torch ops, so the faithful CPU baseline is the same sequence of
torch CPU operators it issues, on the same flat tensors, without the model / autograd around
them.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this module; the product package never does.

Parity status: PINNED.  `tests/test_oracle_golden.py::test_torch_port_matches_goldens` checks
these functions against the fixtures written by the unmodified reference
(`tests/golden/make_golden.py`) and against the numpy oracle.

Operator sequence restated (reference file:line):
  SNIP       train.py:260  |g|            train.py:289  |w| * |g|        train.py:294  cat
             train.py:299  k = int(N*s)   train.py:306-307 full sort, [k-1].item()
             train.py:316  (score > thr).float()
  magnitude  torch/nn/utils/prune.py:1114-1123 two parameters_to_vector, :368-370 slice of the
             alive entries, :526 k = round(amount * n_alive), :536 topk(|t|, k, largest=False),
             :538 scatter of zeros, :410 write-back, :1149-1161 per-tensor slices
"""
import torch


def snip_mask_build(weights, grads_per_batch, target_sparsity):
    """weights: list of fp32 CPU tensors; grads_per_batch: list (batch) of lists of tensors.
    Returns (masks as fp32 tensors, threshold as Python number)."""
    acc = None
    for grads in grads_per_batch:
        g_abs = [g.detach().clone().abs() for g in grads]
        part = [w.abs() * ga for w, ga in zip(weights, g_abs)]
        acc = part if acc is None else [a.add_(p) for a, p in zip(acc, part)]
    flat = torch.cat([s.view(-1) for s in acc])
    n = flat.numel()
    k = int(n * target_sparsity)
    if k >= n:
        thr = float("inf")
    elif k <= 0:
        thr = -1
    else:
        ordered, _ = torch.sort(flat)
        thr = ordered[k - 1].item()
    return [(s > thr).float() for s in acc], thr


def magnitude_mask_build(weights, old_masks, amount):
    """Global L1 pruning of the `amount` fraction (or count) of the surviving entries.
    Returns (masks as fp32 tensors, k)."""
    eff = weights if old_masks is None else [w * m for w, m in zip(weights, old_masks)]
    flat_w = torch.cat([e.reshape(-1) for e in eff])
    flat_m = torch.cat([torch.ones_like(w).reshape(-1) if old_masks is None else old_masks[i].reshape(-1)
                        for i, w in enumerate(weights)])
    alive = flat_m == 1
    t = flat_w[alive]
    n_alive = t.numel()
    k = amount if isinstance(amount, int) else round(amount * n_alive)
    part = torch.ones_like(t)
    if k > 0:
        idx = torch.topk(torch.abs(t).view(-1), k=k, largest=False).indices
        part.view(-1)[idx] = 0
    new_flat = flat_m.clone()
    new_flat[alive] = part
    out, ptr = [], 0
    for w in weights:
        out.append(new_flat[ptr:ptr + w.numel()].view_as(w))
        ptr += w.numel()
    return out, k


def sparsity_percent(weights, masks):
    """train.py:347-369 over effective weights."""
    total = zeros = 0
    for w, m in zip(weights, masks):
        eff = w * m
        total += eff.nelement()
        zeros += torch.sum(eff == 0).item()
    return 0.0 if total == 0 else 100.0 * zeros / total
