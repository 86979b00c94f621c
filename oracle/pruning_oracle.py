"""CPU oracle for the pruning hot path — TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's algorithm (EIDOSLAB/pruning-for-vision-representation
`train.py` + the `torch.nn.utils.prune` arithmetic it calls, PyTorch 2.11.0).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module; the product package never does.

Parity status: PINNED.  `tests/golden/make_golden.py` ran the unmodified reference functions
(`train.snip_pruning`, `train.magnitude_pruning`, `train.compute_sparsity_global`, imported from
/root/reference) in the build container and stored their outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks every function below against those fixtures.

Tie policy: SNIP needs none (strict `>`).  Magnitude pruning removes exactly k
alive entries; among entries equal to the k-th value the reference's choice is an artefact of
`torch.topk` internals, ours is "lowest flat index first" (flat order = named_modules() order,
row-major).  `magnitude_masks` implements that policy; `masks_equal_modulo_ties` is the
comparison the parity tests use against the un-pinned reference.
"""
import numpy as np

F32 = np.float32


# ---- scores ------------------------------------------------------------------------------
def snip_score(w, g):
    """train.py:260 (`grad.abs()`) and train.py:289 (`param.abs() * grads[key]`), fp32."""
    return (np.abs(np.asarray(w, F32)) * np.abs(np.asarray(g, F32))).astype(F32)


def snip_score_accumulate(acc, w, g):
    """Multi-batch extension: acc_b = acc_{b-1} + |w * g_b|, fp32, batch order."""
    s = snip_score(w, g)
    return s if acc is None else (np.asarray(acc, F32) + s).astype(F32)


# ---- threshold ---------------------------------------------------------------------------
def snip_k(numel, target_sparsity):
    """train.py:299: k = int(N * target_sparsity) (double multiply, truncation)."""
    return int(numel * target_sparsity)


def kth_smallest(flat, k):
    """k-th smallest (1-based) with NaN last — sorted(all)[k-1] of train.py:306-307."""
    flat = np.asarray(flat, F32).reshape(-1)
    return F32(np.partition(flat, k - 1)[k - 1])


def snip_threshold(all_scores, target_sparsity):
    """train.py:299-307, returns a Python float like `.item()`."""
    n = all_scores.size
    k = snip_k(n, target_sparsity)
    if k >= n:
        return float("inf")
    if k <= 0:
        return -1
    return float(kth_smallest(all_scores, k))


def snip_masks(scores, threshold):
    """train.py:316: mask = (score > threshold).float() — every tie pruned, NaN pruned."""
    thr = F32(threshold)
    with np.errstate(invalid="ignore"):
        return [(np.asarray(s, F32) > thr) for s in scores]


def snip_pruning(weights, grads_per_batch, target_sparsity):
    """Whole SNIP mask build: list of weights, list (batches) of lists of grads."""
    scores = [None] * len(weights)
    for grads in grads_per_batch:
        scores = [snip_score_accumulate(a, w, g) for a, w, g in zip(scores, weights, grads)]
    flat = np.concatenate([s.reshape(-1) for s in scores])       # train.py:294
    thr = snip_threshold(flat, target_sparsity)
    return snip_masks(scores, thr), thr, scores


# ---- global magnitude ---------------------------------------------------------------------
def magnitude_k(amount, n_alive):
    """torch/nn/utils/prune.py:1331-1354 `_compute_nparams_toprune`: int amount as is, float amount
    -> round(amount * n) with Python's banker's rounding."""
    if isinstance(amount, (int, np.integer)) and not isinstance(amount, bool):
        return int(amount)
    return round(amount * n_alive)


def magnitude_masks(weights, old_masks, amount):
    """prune.global_unstructured + L1Unstructured (prune.py:1038-1161, 321-416, 503-540):
    among alive entries (old mask == 1) prune the k = round(amount * n_alive) smallest |w|;
    dead entries stay dead.  Ties at the k-th value: lowest flat index first.
    Returns (new_masks, info)."""
    flat_w = np.concatenate([np.asarray(w, F32).reshape(-1) for w in weights])
    if old_masks is None:
        flat_m = np.ones(flat_w.size, dtype=bool)
    else:
        flat_m = np.concatenate([np.asarray(m).reshape(-1).astype(bool) for m in old_masks])
    # prune.py:1114: importance scores are module.weight == mask * orig
    eff = np.where(flat_m, flat_w, F32(0))
    alive_idx = np.flatnonzero(flat_m)                            # prune.py:368-370
    n_alive = alive_idx.size
    k = magnitude_k(amount, n_alive)
    new_flat = flat_m.copy()
    info = {"k": k, "n_alive": int(n_alive), "threshold": None, "n_less": 0, "n_equal": 0, "quota": 0}
    if k > 0:
        keys = np.abs(eff[alive_idx])                             # prune.py:536 topk(abs(t), k, largest=False)
        thr = kth_smallest(keys, k)
        with np.errstate(invalid="ignore"):
            if np.isnan(thr):
                less = ~np.isnan(keys)
                equal = np.isnan(keys)
            else:
                less = keys < thr
                equal = keys == thr
        n_less = int(less.sum())
        quota = k - n_less
        tie_idx = alive_idx[equal][:quota]                        # lowest flat index first
        new_flat[alive_idx[less]] = False
        new_flat[tie_idx] = False
        info.update(threshold=float(thr), n_less=n_less, n_equal=int(equal.sum()), quota=int(quota))
    out, ptr = [], 0
    for w in weights:                                             # prune.py:1149-1161
        n = np.asarray(w).size
        out.append(new_flat[ptr:ptr + n].reshape(np.asarray(w).shape))
        ptr += n
    return out, info


def masks_equal_modulo_ties(masks_a, masks_b, weights, threshold):
    """True when two mask sets agree everywhere except (possibly) on entries with
    |w| == threshold, and prune the same NUMBER of those entries."""
    a = np.concatenate([np.asarray(m).reshape(-1).astype(bool) for m in masks_a])
    b = np.concatenate([np.asarray(m).reshape(-1).astype(bool) for m in masks_b])
    w = np.abs(np.concatenate([np.asarray(x, F32).reshape(-1) for x in weights]))
    tied = w == F32(threshold)
    return bool(np.array_equal(a[~tied], b[~tied]) and a[tied].sum() == b[tied].sum())


# ---- sparsity ---------------------------------------------------------------------------
def compute_sparsity_global(weights, masks=None):
    """train.py:347-369: 100 * #(module.weight == 0) / N over the effective (masked) weights."""
    total = zeros = 0
    for i, w in enumerate(weights):
        w = np.asarray(w, F32)
        eff = w if masks is None else np.where(np.asarray(masks[i]).astype(bool), w, F32(0))
        total += eff.size
        zeros += int((eff == 0).sum())
    return 0.0 if total == 0 else 100.0 * zeros / total


# ---- masked SGD step ---------------------------------------------------------------------
def masked_sgd_step(w, g, buf, mask, lr, momentum=0.0, dampening=0.0, weight_decay=0.0,
                    nesterov=False, first_step=False):
    """One optimizer step on `weight_orig` under the prune reparametrisation.
    grad(weight_orig) = grad(weight) * mask              (MulBackward of prune.py:71-74)
    torch/optim/sgd.py:343-380 (_single_tensor_sgd):  g += wd*p; buf = g (first) or mu*buf +
    (1-damp)*g; g = g + mu*buf if nesterov else buf; p -= lr*g.
    Returns (w_new, buf_new, w_eff) with w_eff = mask * w_new (the next forward's weight).
    fp32 throughout, operations unfused (the CUDA path may fuse multiply-adds: compare with a
    tolerance)."""
    w = np.asarray(w, F32); g = np.asarray(g, F32)
    m = np.ones_like(w, dtype=bool) if mask is None else np.asarray(mask).astype(bool)
    g = np.where(m, g, F32(0)).astype(F32)
    if weight_decay != 0:
        g = (g + F32(weight_decay) * w).astype(F32)
    if momentum != 0:
        if first_step or buf is None:
            buf = g.copy()
        else:
            buf = (np.asarray(buf, F32) * F32(momentum)).astype(F32)
            buf = (buf + F32(1 - dampening) * g).astype(F32)
        g = (g + F32(momentum) * buf).astype(F32) if nesterov else buf
    w_new = (w + F32(-lr) * g).astype(F32)
    return w_new, buf, np.where(m, w_new, F32(0)).astype(F32)
