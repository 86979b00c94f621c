"""BASELINE.json configurations at their FULL sizes, checked through size-independent properties
(the oracle is too slow there): rank identities of the threshold verified with independent torch
reductions on the GPU, nesting / exact counts of iterative masks, fused == streaming score passes,
masked-step invariants.  Complements the bit-exact small-size parity tests."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from pruning_for_vision_representation_b200 import _lib as L                      # noqa: E402
from pruning_for_vision_representation_b200.plan import ParamPlan                 # noqa: E402
from pruning_for_vision_representation_b200.shapes import prunable_numels        # noqa: E402

DEV = torch.device("cuda:0")


def views(flat, numels):
    out, o = [], 0
    for n in numels:
        out.append(flat[o:o + n]); o += n
    return out


def unpack_bits(plan, mask, numels):
    """packed words -> flat bool tensor on the GPU (chunk-major layout of include/b200prune.h)."""
    words = mask.view(torch.int32)
    bits = ((words.unsqueeze(1) >> torch.arange(32, device=DEV, dtype=torch.int32)) & 1).bool().reshape(-1)
    out, chunk = [], 0
    for n in numels:
        nchunks = (n + L.CHUNK - 1) // L.CHUNK
        out.append(bits[chunk * L.CHUNK:chunk * L.CHUNK + n]); chunk += nchunks
    return torch.cat(out)


def default_init_like(numels, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    flat = torch.empty(sum(numels), device=DEV)
    for v in views(flat, numels):
        bound = 1.0 / (v.numel() ** 0.25)                 # spread of scales across tensors, like per-layer fan-in inits
        v.uniform_(-bound, bound, generator=g)
    return flat


@pytest.mark.parametrize("model", ["resnet152", "vit_l_16"])
def test_config5_threshold_sweep_rank_identities(model):
    """Config 5: one-shot global magnitude threshold at 0.5/0.8/0.9/0.95/0.99 over ResNet-152 (60.0 M)
    and ViT-L/16 (228.3 M).  For every sparsity: n_less < k <= n_less + n_equal with the counts recomputed
    by torch, exactly k entries pruned, everything below the threshold pruned, everything above kept."""
    numels = prunable_numels(model)
    n = sum(numels)
    assert n == {"resnet152": 60_040_384, "vit_l_16": 228_302_848}[model]
    w = default_init_like(numels, 1)
    absw = w.abs()
    for impl in ("sampled", "exact"):
        plan = ParamPlan(numels, DEV).set_select_impl(impl)
        plan.bind(L.SLOT_W, views(w, numels))
        for s in (0.5, 0.8, 0.9, 0.95, 0.99):
            k = round(s * n)
            mask = plan.new_mask()
            plan.select_kth(L.KEY_ABS_W, k, L.MODE_EXACT_K)
            plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, mask)
            res = plan.result()
            thr = torch.tensor(res["threshold"], device=DEV)
            n_less = int((absw < thr).sum()); n_equal = int((absw == thr).sum())
            assert (res["n_less"], res["n_equal"]) == (n_less, n_equal), (model, impl, s)
            assert n_less < k <= n_less + n_equal and res["quota"] == k - n_less
            assert res["n_kept"] == n - k
            kept = unpack_bits(plan, mask, numels)
            assert int(kept.sum()) == n - k
            assert not kept[absw < thr].any() and kept[absw > thr].all()
            if impl == "sampled":
                assert res["passes_full"] == 1, "the sample bracket should hold on these distributions"
        plan.close()


def test_config2_resnet50_snip_8_batches_full_size():
    """Config 2 at full size: fused multi-batch pass == 8 streaming passes (bit for bit), the mask is
    exactly (score > threshold), and the threshold is the k-th smallest score."""
    numels = prunable_numels("resnet50")
    n = sum(numels); k = int(n * 0.9)
    assert (n, k) == (25_502_912, 22_952_620)
    w = default_init_like(numels, 1)
    grads = [torch.randn(n, device=DEV, generator=torch.Generator(device=DEV).manual_seed(300 + b)) * 1e-3 for b in range(8)]
    plan = ParamPlan(numels, DEV)
    s_stream = torch.empty(n, device=DEV); s_fused = torch.empty(n, device=DEV)
    plan.bind(L.SLOT_W, views(w, numels)).bind(L.SLOT_SCORE, views(s_stream, numels))
    for b, g in enumerate(grads):
        plan.bind(L.SLOT_G, views(g, numels)); plan.score_accumulate(b > 0)
    plan.bind(L.SLOT_SCORE, views(s_fused, numels))
    plan.score_accumulate_multi([plan.pointer_table(L.SLOT_G, views(g, numels)) for g in grads])
    assert torch.equal(s_stream, s_fused)
    ref = torch.zeros(n, device=DEV)
    for g in grads:
        ref += (w * g).abs()
    assert torch.equal(ref, s_fused)                      # same fp32 operations in the same order
    mask = plan.new_mask()
    plan.select_kth(L.KEY_SCORE, k, L.MODE_SNIP_STRICT)
    plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, mask)
    res = plan.result()
    thr = torch.tensor(res["threshold"], device=DEV)
    assert int((s_fused < thr).sum()) < k <= int((s_fused <= thr).sum())
    kept = unpack_bits(plan, mask, numels)
    assert torch.equal(kept, s_fused > thr) and res["n_kept"] == int(kept.sum())
    # the default sequence of bench.py: score pass fused with the select's sweep, emit by patching
    s_sweep = torch.full((n,), -1.0, device=DEV)
    plan.bind(L.SLOT_SCORE, views(s_sweep, numels))
    mask2 = plan.new_mask()
    plan.snip_mask_build([plan.pointer_table(L.SLOT_G, views(g, numels)) for g in grads], k, mask2)
    res2 = plan.result()
    assert torch.equal(s_sweep, s_fused) and torch.equal(mask2, mask)
    assert {key: res2[key] for key in ("k", "n_less", "n_equal", "n_kept", "threshold")} == {key: res[key] for key in ("k", "n_less", "n_equal", "n_kept", "threshold")}
    assert res2["passes_full"] == 1, "the sampled bracket should hold on this distribution"
    # every sparsity of config 5's sweep through the same sequence
    for s in (0.5, 0.8, 0.95, 0.99):
        ks = int(n * s)
        m = plan.new_mask(); plan.snip_mask_build([plan.pointer_table(L.SLOT_G, views(g, numels)) for g in grads], ks, m)
        r = plan.result()
        t = torch.tensor(r["threshold"], device=DEV)
        assert int((s_fused < t).sum()) < ks <= int((s_fused <= t).sum()), s
        assert torch.equal(unpack_bits(plan, m, numels), s_fused > t), s


def test_config4_vit_b16_iterative_pruning_with_masked_steps():
    """Config 4: ViT-B/16 prunable set (65.06 M), 14 rounds of 20 % of the survivors
    (20 -> 36 -> 48.8 -> ... -> 95.6 %, train.py:656-708) with masked SGD steps (bf16 weight emit) between."""
    numels = prunable_numels("vit_b_16")
    n = sum(numels)
    assert n == 65_058_816
    w = default_init_like(numels, 2)
    buf = torch.zeros(n, device=DEV); weff16 = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    plan = ParamPlan(numels, DEV)
    plan.bind(L.SLOT_W, views(w, numels)).bind(L.SLOT_BUF, views(buf, numels)).bind(L.SLOT_WEFF16, views(weff16, numels))
    old, n_alive, prev_kept = None, n, None
    gen = torch.Generator(device=DEV).manual_seed(7)
    for r in range(14):
        k = round(0.2 * n_alive)                           # torch/nn/utils/prune.py:1331-1354
        new = plan.new_mask()
        plan.select_kth(L.KEY_ABS_W, k, L.MODE_EXACT_K, old)
        plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, new, old)
        res = plan.result()
        n_alive -= k
        assert res["n_valid"] == n_alive + k and res["n_kept"] == n_alive
        kept = unpack_bits(plan, new, numels)
        assert int(kept.sum()) == n_alive
        if prev_kept is not None:
            assert not (kept & ~prev_kept).any()           # masks only shrink: a pruned weight never comes back
            alive_w = w.abs()[prev_kept]
            thr = torch.tensor(res["threshold"], device=DEV)
            assert int((alive_w < thr).sum()) == res["n_less"] and int((alive_w == thr).sum()) == res["n_equal"]
        zeros, bits = plan.count_zeros(new, use_weights=False)
        assert bits == n_alive and abs(100.0 * zeros / n - 100.0 * (1 - 0.8 ** (r + 1))) < 1e-3
        # two masked SGD steps with synthetic gradients: pruned entries never reach the forward weight
        for step in range(2):
            g = torch.randn(n, device=DEV, generator=gen) * 1e-3
            plan.bind(L.SLOT_G, views(g, numels))
            plan.masked_sgd_step(new, 0.1, 0.9, 0.0, 1e-4, L.SGD_EMIT_WEFF16 | (L.SGD_FIRST_STEP if (r == 0 and step == 0) else 0))
        assert int(torch.count_nonzero(weff16[~kept])) == 0
        assert torch.equal(weff16[kept], w[kept].to(torch.bfloat16))
        old, prev_kept = new, kept
    assert abs(100.0 * (n - n_alive) / n - 95.60) < 0.01     # the loop of train.py:666 stops after this round


def test_config3_lost_batch_256_deterministic_and_sane():
    from pruning_for_vision_representation_b200 import object_discovery as OD
    feats = torch.randn(256, 900, 384, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    a = OD.lost_batched(feats, [30, 30], [16, 16], (3, 480, 480))
    b = OD.lost_batched(feats, [30, 30], [16, 16], (3, 480, 480))
    assert torch.equal(a["box"], b["box"]) and torch.equal(a["seed"], b["seed"])       # run-to-run identical
    box, seed, status = a["box"].cpu().numpy(), a["seed"].cpu().numpy(), a["status"].cpu().numpy()
    assert (status == 0).all()                             # the seed is always inside its own component
    assert (box[:, 0] < box[:, 2]).all() and (box[:, 1] < box[:, 3]).all() and box.min() >= 0 and box.max() <= 480
    sx, sy = (seed % 30) * 16, (seed // 30) * 16            # the seed patch lies inside its box
    assert ((box[:, 0] <= sx) & (sx < box[:, 2]) & (box[:, 1] <= sy) & (sy < box[:, 3])).all()
    deg = torch.stack(list(a["degree"]))
    A0 = feats[0] @ feats[0].T
    ref_deg = ((A0 > 0).sum(dim=1) - 1).to(torch.int32)     # minus the diagonal
    assert int((deg[0] != ref_deg).sum()) <= 2              # sign of near-zero entries may differ between summation orders
    assert seed[0] == int(torch.argmin(deg[0]))
