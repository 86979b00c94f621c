"""GPU parity of the pruning kernels against the CPU oracle and the reference goldens.
Everything goes through the C-ABI (ctypes) — see pruning_for_vision_representation_b200/plan.py."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import pruning_oracle as PO

pytestmark = pytest.mark.gpu

from pruning_for_vision_representation_b200 import _lib as L           # noqa: E402
from pruning_for_vision_representation_b200.plan import ParamPlan, pack_mask_words   # noqa: E402

DEV = "cuda:0"


def _seq(z, prefix):
    out, i = [], 0
    while f"{prefix}{i}" in z:
        out.append(z[f"{prefix}{i}"])
        i += 1
    return out


def to_dev(arrs, misalign=False, dtype=torch.float32):
    """numpy arrays -> contiguous CUDA tensors; misalign=True places them 4 bytes off a 16-B boundary."""
    out = []
    for a in arrs:
        a = np.asarray(a)
        t = torch.from_numpy(a.reshape(-1).copy())
        if misalign:
            big = torch.empty(t.numel() + 8, dtype=t.dtype, device=DEV)
            off = 1 if big.data_ptr() % 16 == 0 else 0
            v = big[off:off + t.numel()]
            assert v.data_ptr() % 16 != 0
            v.copy_(t)
            out.append(v)
        else:
            out.append(t.to(DEV))
    return out


SELECT_IMPL = "sampled"


@pytest.fixture(params=["sampled", "exact"], autouse=True)
def select_impl(request):
    """Every test runs with both select implementations; they must give identical results."""
    global SELECT_IMPL
    SELECT_IMPL = request.param
    yield request.param
    SELECT_IMPL = "sampled"


def make_plan(arrs, **kw):
    return ParamPlan([int(np.asarray(a).size) for a in arrs], DEV, **kw).set_select_impl(SELECT_IMPL)


def gpu_masks(plan, mask):
    return plan.unpack_mask_host(mask)


def run_magnitude(plan, k, old_mask=None):
    new = plan.new_mask()
    if k == 0:
        plan.select_begin(0, L.MODE_EXACT_K)
        plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, new, old_mask, force=1)
    else:
        plan.select_kth(L.KEY_ABS_W, k, L.MODE_EXACT_K, old_mask)
        plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, new, old_mask)
    return new, plan.result()


@pytest.mark.parametrize("misalign", [False, True])
def test_score_accumulate_bit_exact(misalign):
    rng = np.random.default_rng(1)
    sizes = [351, 2808, 25728, 670, 4096, 8192 + 4, 1]
    w = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    plan = make_plan(w)
    wt = to_dev(w, misalign)
    st = [torch.full((n,), 7.0, device=DEV) for n in sizes]           # garbage: assign must overwrite
    plan.bind(L.SLOT_W, wt).bind(L.SLOT_SCORE, st)
    acc = [None] * len(sizes)
    for b in range(3):
        g = [(1e-3 * rng.standard_normal(n)).astype(np.float32) for n in sizes]
        if b == 1:
            g[2][5] = np.nan; g[0][3] = np.inf; w[0][3] = 0.0          # inf * 0 = NaN, NaN propagates
            wt[0][3] = 0.0
        gt = to_dev(g, misalign)
        plan.bind(L.SLOT_G, gt)
        plan.score_accumulate(accumulate=(b > 0))
        acc = [PO.snip_score_accumulate(a, ww, gg) for a, ww, gg in zip(acc, w, g)]
        torch.cuda.synchronize()
        for s, a in zip(st, acc):
            got = s.cpu().numpy()
            assert np.array_equal(got, a, equal_nan=True)
            assert not np.signbit(got[~np.isnan(got)]).any()


@pytest.mark.parametrize("misalign", [False, True])
def test_score_multi_is_bit_identical_to_sequential(misalign):
    """One fused pass over B gradient sets == B successive accumulate launches, bit for bit."""
    rng = np.random.default_rng(7)
    sizes = [351, 2808, 25728, 4096 * 2, 8192 + 4, 1]
    w = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    plan = make_plan(w)
    plan.bind(L.SLOT_W, to_dev(w, misalign))
    for n_sets in (1, 2, 3, 5, 8, 11):
        grads = [[(1e-3 * rng.standard_normal(n)).astype(np.float32) for n in sizes] for _ in range(n_sets)]
        gdev = [to_dev(g, misalign) for g in grads]
        seq = [torch.full((n,), 3.0, device=DEV) for n in sizes]
        plan.bind(L.SLOT_SCORE, seq)
        for b in range(n_sets):
            plan.bind(L.SLOT_G, gdev[b]); plan.score_accumulate(b > 0)
        fused = [torch.full((n,), -7.0, device=DEV) for n in sizes]
        plan.bind(L.SLOT_SCORE, fused)
        tables = [plan.pointer_table(L.SLOT_G, g) for g in gdev]
        plan.score_accumulate_multi(tables, accumulate=False)
        for a, b_ in zip(seq, fused):
            assert torch.equal(a, b_), n_sets
        # accumulate=True continues an existing score
        m = min(2, n_sets)
        plan.score_accumulate_multi(tables[:m], accumulate=True)
        plan.bind(L.SLOT_SCORE, seq)
        for b in range(m):
            plan.bind(L.SLOT_G, gdev[b]); plan.score_accumulate(True)
        for a, b_ in zip(seq, fused):
            assert torch.equal(a, b_), n_sets
        acc = [None] * len(sizes)
        for g in grads + grads[:m]:
            acc = [PO.snip_score_accumulate(x, ww, gg) for x, ww, gg in zip(acc, w, g)]
        for a, e in zip(fused, acc):
            assert np.array_equal(a.cpu().numpy(), e)


def test_snip_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "snip_tiny.npz"))
    w, g, ref = _seq(z, "w"), _seq(z, "g"), _seq(z, "m")
    plan = make_plan(w)
    st = [torch.empty(a.size, device=DEV) for a in w]
    plan.bind(L.SLOT_W, to_dev(w)).bind(L.SLOT_G, to_dev(g)).bind(L.SLOT_SCORE, st)
    plan.score_accumulate(False)
    k = PO.snip_k(plan.total, 0.9)
    plan.select_kth(L.KEY_SCORE, k, L.MODE_SNIP_STRICT)
    mask = plan.new_mask()
    plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, mask)
    res = plan.result()
    assert float(np.float32(res["threshold"])) == float(z["threshold"][0])
    for m, r in zip(gpu_masks(plan, mask), ref):
        assert np.array_equal(m, r.reshape(-1).astype(bool))
    assert res["n_kept"] == sum(int(r.sum()) for r in ref)


@pytest.mark.parametrize("misalign", [False, True])
def test_magnitude_golden_iterative(golden_dir, misalign):
    z = np.load(os.path.join(golden_dir, "magnitude_tiefree.npz"))
    w = _seq(z, "w")
    plan = make_plan(w)
    plan.bind(L.SLOT_W, to_dev(w, misalign))
    old, n_alive = None, plan.total
    for r, amount in enumerate(z["amounts"]):
        k = PO.magnitude_k(float(amount), n_alive)
        new, res = run_magnitude(plan, k, old)
        ref = _seq(z, f"m{r}_")
        got = gpu_masks(plan, new)
        if res["quota"] == res["n_equal"]:
            for m, rm in zip(got, ref):
                assert np.array_equal(m, rm.reshape(-1).astype(bool)), f"round {r}"
        else:
            assert PO.masks_equal_modulo_ties(got, ref, w, res["threshold"])
        assert res["n_valid"] == n_alive and res["n_kept"] == n_alive - k
        zeros, bits = plan.count_zeros(new, use_weights=True)
        assert 100.0 * zeros / plan.total == float(z["sparsity"][r])
        assert bits == res["n_kept"]
        # continue from the reference's masks (packed on the host)
        old = torch.from_numpy(pack_mask_words([rm.reshape(-1) for rm in ref], plan.numels).view(np.int32)).to(DEV)
        n_alive = sum(int(rm.sum()) for rm in ref)


def test_magnitude_planted_ties_policy(golden_dir):
    z = np.load(os.path.join(golden_dir, "magnitude_tiny.npz"))
    w = _seq(z, "w")
    plan = make_plan(w)
    plan.bind(L.SLOT_W, to_dev(w))
    k = PO.magnitude_k(0.5, plan.total)
    new, res = run_magnitude(plan, k)
    exp, info = PO.magnitude_masks(w, None, 0.5)
    assert (res["n_less"], res["n_equal"], res["quota"]) == (info["n_less"], info["n_equal"], info["quota"])
    assert res["quota"] < res["n_equal"]                       # the tie path really ran
    for m, e in zip(gpu_masks(plan, new), exp):
        assert np.array_equal(m, e.reshape(-1))               # lowest-flat-index-first, bit exact vs oracle
    assert PO.masks_equal_modulo_ties(gpu_masks(plan, new), _seq(z, "m1_"), w, res["threshold"])
    assert res["n_kept"] == plan.total - k


@pytest.mark.parametrize("n_ties", [3, 40, 511, 512, 513, 3000])
def test_magnitude_tie_list_and_table_paths(n_ties):
    """A planted tied set at the cut: up to 512 ties go through the short position list (tie_list_pick), more through the
    per-chunk table (tie_scan_body); both must drop the lowest flat indices first, bit exact vs the oracle, wherever the
    quota cuts the set (prune.py:1145-1153 semantics pinned by SURVEY 8c)."""
    rng = np.random.default_rng(11 + n_ties)
    sizes = [4096 * 9 + 17, 300, 4096 * 30, 4096 * 4 + 4095]
    total = sum(sizes)
    flat = (rng.standard_normal(total) * 0.05).astype(np.float32)
    tie_val = np.float32(0.0337)
    pos = np.sort(rng.choice(total, n_ties, replace=False))
    flat[pos] = tie_val * rng.choice(np.array([1.0, -1.0], np.float32), n_ties)
    w, o = [], 0
    for n in sizes:
        w.append(flat[o:o + n].copy()); o += n
    plan = make_plan(w)
    plan.bind(L.SLOT_W, to_dev(w))
    n_less = int((np.abs(flat) < tie_val).sum())
    for quota in sorted({1, 2, n_ties // 2, n_ties - 1}):
        if not 0 < quota < n_ties:
            continue
        k = n_less + quota
        new, res = run_magnitude(plan, k)
        exp, info = PO.magnitude_masks(w, None, int(k))
        assert (res["n_less"], res["n_equal"], res["quota"]) == (info["n_less"], info["n_equal"], info["quota"]) == (n_less, n_ties, quota)
        for m, e in zip(gpu_masks(plan, new), exp):
            assert np.array_equal(m, e.reshape(-1)), (n_ties, quota)
        assert res["n_kept"] == total - k


@pytest.mark.parametrize("cand_capacity", [0, 64])
@pytest.mark.parametrize("case", ["random", "constant", "few_values", "nan_tail", "denormal"])
def test_select_edge_cases(case, cand_capacity):
    rng = np.random.default_rng(3)
    sizes = [5000, 4096, 12289, 33]
    if case == "random":
        w = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    elif case == "constant":
        w = [np.full(n, -0.25, np.float32) for n in sizes]
    elif case == "few_values":
        w = [rng.choice(np.array([0.0, -0.0, 0.5, -0.5, 2.0], np.float32), n) for n in sizes]
    elif case == "nan_tail":
        w = [rng.standard_normal(n).astype(np.float32) for n in sizes]
        w[1][::7] = np.nan
    else:
        w = [(rng.standard_normal(n) * 1e-41).astype(np.float32) for n in sizes]
    plan = make_plan(w, cand_capacity=cand_capacity)
    plan.bind(L.SLOT_W, to_dev(w))
    total = plan.total
    for k in (1, 2, total // 3, total // 2, total - 1, total):
        new, res = run_magnitude(plan, k)
        exp, info = PO.magnitude_masks(w, None, int(k))
        assert (res["n_less"], res["n_equal"], res["quota"]) == (info["n_less"], info["n_equal"], info["quota"]), (case, k)
        if not np.isnan(info["threshold"]):
            assert float(np.float32(res["threshold"])) == info["threshold"]
        for m, e in zip(gpu_masks(plan, new), exp):
            assert np.array_equal(m, e.reshape(-1)), (case, k)
        assert res["n_kept"] == total - k
        if cand_capacity == 64 and case in ("constant", "few_values"):
            # bucket larger than the candidate buffer: the exact select runs 3 full passes (histogram mode);
            # the sampled select detects the overflow in its sweep and falls back to those 3 passes
            assert res["passes_full"] == (3 if SELECT_IMPL == "exact" else 4)


@pytest.mark.parametrize("misalign", [False, True])
@pytest.mark.parametrize("n_sets", [1, 3, 8, 9])
def test_snip_mask_build_refresh_repoints_the_tables(misalign, n_sets):
    """b200p_snip_mask_build_refresh: the tables still point at LAST build's gradient tensors (overwritten with garbage
    here); the sample kernel re-points them at the new ones itself.  Same scores, mask and result block as
    update_tables + snip_mask_build; 9 sets take the documented fallback (table update + build)."""
    rng = np.random.default_rng(500 + n_sets)
    sizes = [500_001, 4096 * 21, 77_777, 33, 4096]
    w = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    plan = make_plan(w)
    plan.bind(L.SLOT_W, to_dev(w, misalign))
    stale = [to_dev([np.full(n, 1e30, np.float32) for n in sizes]) for _ in range(n_sets)]
    tables = [plan.pointer_table(L.SLOT_G, st) for st in stale]
    for build in range(2):                                               # two builds: fresh tensors each time
        g = [[(1e-3 * rng.standard_normal(n)).astype(np.float32) for n in sizes] for _ in range(n_sets)]
        gd = [to_dev(gb, misalign) for gb in g]
        k = PO.snip_k(plan.total, 0.9 if build == 0 else 0.5)
        s_new = [torch.full((n,), -3.0, device=DEV) for n in sizes]
        plan.bind(L.SLOT_SCORE, s_new)
        m_new = plan.new_mask(); plan.snip_mask_build_refresh(tables, gd, k, m_new)
        r_new = plan.result()
        # reference: explicit table update, then the plain fused build, on a second plan
        ref = make_plan(w)
        ref.bind(L.SLOT_W, to_dev(w, misalign))
        s_ref = [torch.zeros(n, device=DEV) for n in sizes]
        ref.bind(L.SLOT_SCORE, s_ref)
        rt = [ref.pointer_table(L.SLOT_G, gb) for gb in gd]
        m_ref = ref.new_mask(); ref.snip_mask_build(rt, k, m_ref)
        r_ref = ref.result()
        for a, b in zip(s_new, s_ref):
            assert torch.equal(a, b)
        assert torch.equal(m_new, m_ref)
        for key in ("k", "n_less", "n_equal", "n_kept", "threshold", "thr_key"):
            assert r_new[key] == r_ref[key], key
        exp, thr, _ = PO.snip_pruning(w, g, 0.9 if build == 0 else 0.5)
        for got, e in zip(gpu_masks(plan, m_new), exp):
            assert np.array_equal(got, np.asarray(e).reshape(-1).astype(bool))
        # the tables now point at this build's tensors: the plain build gives the same mask again
        m_again = plan.new_mask(); plan.snip_mask_build(tables, k, m_again)
        assert torch.equal(m_again, m_new)


def test_snip_strict_degenerate():
    """All scores zero -> threshold 0 -> everything pruned (fresh ViT head)."""
    sizes = [4096 * 2, 100]
    s = [np.zeros(n, np.float32) for n in sizes]
    plan = make_plan(s)
    plan.bind(L.SLOT_SCORE, to_dev(s))
    plan.select_kth(L.KEY_SCORE, plan.total // 2, L.MODE_SNIP_STRICT)
    mask = plan.new_mask()
    plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, mask)
    res = plan.result()
    assert res["threshold"] == 0.0 and res["n_kept"] == 0
    assert int(mask.count_nonzero()) == 0
    # forced thresholds of train.py:300-303
    plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, mask, force=3, forced_threshold=-1.0)
    assert plan.result()["n_kept"] == plan.total
    plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, mask, force=3, forced_threshold=float("inf"))
    assert plan.result()["n_kept"] == 0


def test_large_random_select_vs_partition():
    rng = np.random.default_rng(5)
    sizes = [3_000_000, 1_234_567, 2_359_296]
    w = [(rng.standard_normal(n) * s).astype(np.float32) for n, s in zip(sizes, (0.02, 0.05, 0.01))]
    plan = make_plan(w)
    plan.bind(L.SLOT_W, to_dev(w))
    flat = np.abs(np.concatenate(w))
    for sp in (0.5, 0.9, 0.99):
        k = round(sp * plan.total)
        new, res = run_magnitude(plan, k)
        thr = np.partition(flat, k - 1)[k - 1]
        assert np.float32(res["threshold"]) == thr
        assert res["n_less"] == int((flat < thr).sum()) and res["n_equal"] == int((flat == thr).sum())
        assert res["n_kept"] == plan.total - k and res["passes_full"] == (2 if SELECT_IMPL == "exact" else 1)
        got = np.concatenate(gpu_masks(plan, new))
        assert np.array_equal(got[flat != thr], (flat > thr)[flat != thr])


def test_mask_build_equals_select_plus_emit():
    """b200p_mask_build (sweep writes the provisional mask into the destination, emit patches it) gives the
    same packed words and result block as select_kth + emit_masks, with and without an old mask, with ties."""
    rng = np.random.default_rng(21)
    sizes = [700_001, 4096 * 50, 123_457, 33]
    w = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    for a in w:
        idx = rng.choice(a.size, size=a.size // 6, replace=False)
        a[idx] = np.float32(0.3)                                           # a big tied set around the 20-30 % quantile
    plan = make_plan(w)
    plan.bind(L.SLOT_W, to_dev(w))
    sc = to_dev([np.abs(a) * np.float32(1e-3) for a in w])
    plan.bind(L.SLOT_SCORE, sc)
    old = None
    n_alive = plan.total
    for amount in (0.25, 0.2, 0.5):
        k = PO.magnitude_k(amount, n_alive)
        m1 = plan.new_mask(); plan.select_kth(L.KEY_ABS_W, k, L.MODE_EXACT_K, old); plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, m1, old)
        r1 = plan.result()
        m2 = plan.new_mask(); plan.mask_build(L.KEY_ABS_W, k, L.MODE_EXACT_K, m2, old)
        r2 = plan.result()
        assert torch.equal(m1, m2) and r1 == r2
        assert r1["n_kept"] == n_alive - k and int(m1.view(torch.int32).cpu().numpy().view(np.uint32).astype(np.uint64).sum()) >= 0
        zeros, bits = plan.count_zeros(m1, use_weights=False)
        assert bits == n_alive - k
        old, n_alive = m1, n_alive - k
    k = int(plan.total * 0.9)
    m1 = plan.new_mask(); plan.select_kth(L.KEY_SCORE, k, L.MODE_SNIP_STRICT); plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, m1)
    r1 = plan.result()
    m2 = plan.new_mask(); plan.mask_build(L.KEY_SCORE, k, L.MODE_SNIP_STRICT, m2)
    assert torch.equal(m1, m2) and r1 == plan.result()
    flat = np.concatenate([np.abs(a) * np.float32(1e-3) for a in w])
    assert np.array_equal(np.concatenate(gpu_masks(plan, m1)), flat > np.float32(r1["threshold"]))


@pytest.mark.parametrize("misalign", [False, True])
@pytest.mark.parametrize("n_sets", [1, 3, 8, 11])
def test_snip_mask_build_fused_equals_unfused(misalign, n_sets):
    """b200p_snip_mask_build (the pass that writes the scores classifies them against the sampled bracket) gives the
    same scores, packed mask and result block as score_accumulate_multi + mask_build, and the oracle's mask."""
    rng = np.random.default_rng(100 + n_sets)
    sizes = [700_001, 4096 * 37, 123_457, 33, 4096]
    w = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    g = [[(1e-3 * rng.standard_normal(n)).astype(np.float32) for n in sizes] for _ in range(n_sets)]
    g[0][0][17] = np.nan                                                # NaN scores sort last and are pruned
    w[1][5] = 0.0                                                       # exact zero score
    plan = make_plan(w)
    plan.bind(L.SLOT_W, to_dev(w, misalign))
    tables = [plan.pointer_table(L.SLOT_G, to_dev(gb, misalign)) for gb in g]
    acc = [None] * len(sizes)
    for gb in g:
        acc = [PO.snip_score_accumulate(a, ww, gg) for a, ww, gg in zip(acc, w, gb)]
    flat = np.concatenate(acc)
    for sparsity in (0.9, 0.5, 0.01, 0.999):
        k = PO.snip_k(plan.total, sparsity)
        s1 = [torch.zeros(n, device=DEV) for n in sizes]
        plan.bind(L.SLOT_SCORE, s1)
        plan.score_accumulate_multi(tables)
        m1 = plan.new_mask(); plan.mask_build(L.KEY_SCORE, k, L.MODE_SNIP_STRICT, m1)
        r1 = plan.result()
        s2 = [torch.full((n,), 7.0, device=DEV) for n in sizes]
        plan.bind(L.SLOT_SCORE, s2)
        m2 = plan.new_mask(); plan.snip_mask_build(tables, k, m2)
        r2 = plan.result()
        for a, b, ref in zip(s1, s2, acc):
            assert np.array_equal(b.cpu().numpy(), ref, equal_nan=True) and torch.equal(a.isnan(), b.isnan())
        assert torch.equal(m1, m2)
        for key in ("k", "n_less", "n_equal", "n_kept", "threshold", "thr_key"):
            assert r1[key] == r2[key], key
        thr = np.sort(flat)[k - 1]                                       # train.py:306-307 (NaN sorts last)
        assert np.float32(r2["threshold"]) == thr
        assert np.array_equal(np.concatenate(gpu_masks(plan, m2)), flat > thr)          # train.py:316
        # the two-call form with an old mask and fp32 mask outputs
        s3 = [torch.zeros(n, device=DEV) for n in sizes]
        plan.bind(L.SLOT_SCORE, s3)
        plan.snip_score_select(tables, k)
        m3 = plan.new_mask(); plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, m3)
        assert torch.equal(m3, m2)


def test_snip_mask_build_fused_degenerate_scores():
    """Scores the sample cannot bracket (all equal / all zero / a huge tied set at the threshold) take the exact
    fallback inside the finish kernel, over the scores the fused pass has just written."""
    rng = np.random.default_rng(5)
    sizes = [4096 * 9 + 5, 50_000]
    plan = make_plan([np.zeros(n, np.float32) for n in sizes])
    cases = []
    w0 = [np.ones(n, np.float32) for n in sizes]; g0 = [np.full(n, 0.5, np.float32) for n in sizes]; cases.append((w0, g0))    # constant
    cases.append(([np.zeros(n, np.float32) for n in sizes], g0))                                                                # fresh-ViT case: all zero
    w2 = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    g2 = [np.where(rng.random(n) < 0.6, 0.0, 1e-3 * rng.standard_normal(n)).astype(np.float32) for n in sizes]               # 60 % exact zeros
    cases.append((w2, g2))
    for w, g in cases:
        plan.bind(L.SLOT_W, to_dev(w))
        sc = [torch.empty(n, device=DEV) for n in sizes]
        plan.bind(L.SLOT_SCORE, sc)
        tables = [plan.pointer_table(L.SLOT_G, to_dev(g))]
        flat = np.concatenate([PO.snip_score_accumulate(None, ww, gg) for ww, gg in zip(w, g)])
        for sparsity in (0.5, 0.9):
            k = PO.snip_k(plan.total, sparsity)
            m = plan.new_mask(); plan.snip_mask_build(tables, k, m)
            r = plan.result()
            thr = np.sort(flat)[k - 1]
            assert np.float32(r["threshold"]) == thr
            assert np.array_equal(np.concatenate(gpu_masks(plan, m)), flat > thr)
            assert np.array_equal(np.concatenate([t.cpu().numpy() for t in sc]), flat)


def test_snip_mask_build_fused_repeatable():
    """Candidate order in the buffer depends on scheduling, the mask, the scores and the result block must not:
    ten builds of a ResNet-18-sized set, bit-identical every time."""
    rng = np.random.default_rng(77)
    sizes = [2_359_296, 1_179_648, 589_824, 147_456, 36_864, 9_408, 512_000 + 3]
    w = [torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(DEV) for n in sizes]
    plan = make_plan([np.empty(n, np.float32) for n in sizes])
    plan.bind(L.SLOT_W, w)
    tables = [plan.pointer_table(L.SLOT_G, [torch.from_numpy((1e-3 * rng.standard_normal(n)).astype(np.float32)).to(DEV) for n in sizes])
              for _ in range(4)]
    sc = [torch.empty(n, device=DEV) for n in sizes]
    plan.bind(L.SLOT_SCORE, sc)
    k = PO.snip_k(plan.total, 0.9)
    m0 = plan.new_mask(); plan.snip_mask_build(tables, k, m0); r0 = plan.result()
    s0 = torch.cat(sc).clone()
    for _ in range(9):
        for t in sc:
            t.fill_(-1.0)
        m = plan.new_mask(); plan.snip_mask_build(tables, k, m)
        assert torch.equal(m, m0) and plan.result() == r0 and torch.equal(torch.cat(sc), s0)
    assert r0["n_kept"] == plan.total - r0["n_less"] - r0["n_equal"] and r0["n_less"] + 1 <= k <= r0["n_less"] + r0["n_equal"]


def test_mask_roundtrip_apply_and_grads():
    rng = np.random.default_rng(9)
    sizes = [4096 * 3, 777, 4100]
    w = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    masks = [rng.random(n) > 0.7 for n in sizes]
    plan = make_plan(w)
    mf = to_dev([m.astype(np.float32) for m in masks])
    weff = [torch.empty(n, device=DEV) for n in sizes]
    w16 = [torch.empty(n, device=DEV, dtype=torch.bfloat16) for n in sizes]
    g = to_dev([np.ones(n, np.float32) for n in sizes])
    plan.bind(L.SLOT_W, to_dev(w)).bind(L.SLOT_MASKF, mf).bind(L.SLOT_WEFF, weff).bind(L.SLOT_WEFF16, w16).bind(L.SLOT_G, g)
    packed = plan.new_mask()
    plan.mask_pack_from_f32(packed)
    assert np.array_equal(packed.cpu().numpy().view(np.uint32), pack_mask_words(masks, sizes))
    for t in mf:
        t.fill_(5.0)
    plan.mask_unpack_to_f32(packed)
    plan.apply_mask(packed, L.EMIT_WEFF | L.SGD_EMIT_WEFF16)
    plan.mask_grads(packed)
    torch.cuda.synchronize()
    for i, n in enumerate(sizes):
        assert np.array_equal(mf[i].cpu().numpy(), masks[i].astype(np.float32))
        exp = np.where(masks[i], w[i], 0).astype(np.float32)
        assert np.array_equal(weff[i].cpu().numpy(), exp)
        assert torch.equal(w16[i].cpu(), torch.from_numpy(exp).to(torch.bfloat16))
        assert np.array_equal(g[i].cpu().numpy(), masks[i].astype(np.float32))
    zeros, bits = plan.count_zeros(packed, use_weights=False)
    assert bits == sum(int(m.sum()) for m in masks) and zeros == plan.total - bits


@pytest.mark.parametrize("nesterov", [False, True])
def test_masked_sgd_vs_oracle(nesterov):
    rng = np.random.default_rng(13)
    sizes = [4096 * 2, 1000, 4097]
    w = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    masks = [rng.random(n) > 0.5 for n in sizes]
    plan = make_plan(w)
    wt = to_dev(w)
    gt = [torch.empty(n, device=DEV) for n in sizes]
    bt = [torch.zeros(n, device=DEV) for n in sizes]
    weff = [torch.empty(n, device=DEV) for n in sizes]
    w16 = [torch.empty(n, device=DEV, dtype=torch.bfloat16) for n in sizes]
    plan.bind(L.SLOT_W, wt).bind(L.SLOT_G, gt).bind(L.SLOT_BUF, bt).bind(L.SLOT_WEFF, weff).bind(L.SLOT_WEFF16, w16)
    packed = torch.from_numpy(pack_mask_words(masks, sizes).view(np.int32)).to(DEV)
    ow, ob = [x.copy() for x in w], [None] * len(sizes)
    for step in range(4):
        g = [rng.standard_normal(n).astype(np.float32) for n in sizes]
        for t, a in zip(gt, g):
            t.copy_(torch.from_numpy(a))
        flags = L.SGD_EMIT_WEFF | L.SGD_EMIT_WEFF16 | (L.SGD_NESTEROV if nesterov else 0) | (L.SGD_FIRST_STEP if step == 0 else 0)
        plan.masked_sgd_step(packed, 0.1, 0.9, 0.0, 1e-4, flags)
        torch.cuda.synchronize()
        for i in range(len(sizes)):
            ow[i], ob[i], oeff = PO.masked_sgd_step(ow[i], g[i], ob[i], masks[i], 0.1, 0.9, 0.0, 1e-4, nesterov, step == 0)
            np.testing.assert_allclose(wt[i].cpu().numpy(), ow[i], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(bt[i].cpu().numpy(), ob[i], rtol=2e-6, atol=1e-7)
            got_eff = weff[i].cpu().numpy()
            assert np.array_equal(got_eff != 0, masks[i] & (wt[i].cpu().numpy() != 0))     # never re-densifies
            assert np.array_equal(got_eff[masks[i]], wt[i].cpu().numpy()[masks[i]])
            assert torch.equal(w16[i].cpu(), weff[i].cpu().to(torch.bfloat16))
            # pruned entries of weight_orig keep decaying (wd) exactly like the reference
            assert np.all(np.abs(ow[i][~masks[i]]) <= np.abs(w[i][~masks[i]]) + 1e-7)


def test_host_buffer_entry_points(golden_dir):
    z = np.load(os.path.join(golden_dir, "snip_tiny.npz"))
    w, g, ref = _seq(z, "w"), _seq(z, "g"), _seq(z, "m")
    plan = make_plan(w)
    flat_w = torch.from_numpy(np.concatenate([a.reshape(-1) for a in w])).pin_memory()
    rng = np.random.default_rng(2)
    g2 = [(1e-3 * rng.standard_normal(a.shape)).astype(np.float32) for a in w]
    flat_g = [torch.from_numpy(np.concatenate([a.reshape(-1) for a in gs])).pin_memory() for gs in (g, g2)]
    out = torch.zeros(plan.mask_words, dtype=torch.int32).pin_memory()
    # one batch == the reference
    res = plan.snip_mask_build_host(flat_w, flat_g[:1], PO.snip_k(plan.total, 0.9), out)
    assert float(np.float32(res["threshold"])) == float(z["threshold"][0])
    for m, r in zip(plan.unpack_mask_host(out), ref):
        assert np.array_equal(m, r.reshape(-1).astype(bool))
    # two batches == the accumulation oracle
    res = plan.snip_mask_build_host(flat_w, flat_g, PO.snip_k(plan.total, 0.75), out)
    exp, thr, _ = PO.snip_pruning(w, [g, g2], 0.75)
    assert float(np.float32(res["threshold"])) == thr
    for m, e in zip(plan.unpack_mask_host(out), exp):
        assert np.array_equal(m, e.reshape(-1))
    # magnitude from host buffers, with an old mask
    exp1, info1 = PO.magnitude_masks(w, None, 0.3)
    res = plan.magnitude_mask_build_host(flat_w, info1["k"], out)
    for m, e in zip(plan.unpack_mask_host(out), exp1):
        assert np.array_equal(m, e.reshape(-1))
    old = out.clone().pin_memory()
    exp2, info2 = PO.magnitude_masks(w, exp1, 0.2)
    res = plan.magnitude_mask_build_host(flat_w, info2["k"], out, old)
    for m, e in zip(plan.unpack_mask_host(out), exp2):
        assert np.array_equal(m, e.reshape(-1))
    assert res["n_valid"] == info2["n_alive"]


def test_resnet18_known_answer(golden_dir):
    """Config 1: ResNet-18, seed 1, global magnitude 50 % then 20 % of survivors — mask bits
    identical to the reference's (sha256 over the flat mask), modulo the recorded tied entries."""
    import torchvision
    info = json.load(open(os.path.join(golden_dir, "resnet18_magnitude.json")))
    torch.manual_seed(1)
    model = torchvision.models.get_model("resnet18", weights=None, num_classes=1000)
    ws = [m.weight.detach() for m in model.modules() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
    plan = ParamPlan([t.numel() for t in ws], DEV)
    assert plan.total == info["N"]
    plan.bind(L.SLOT_W, [t.reshape(-1).to(DEV) for t in ws])

    def digest(mask, tied_idx, tied_ref):
        flat = np.concatenate(plan.unpack_mask_host(mask))
        for i, v in zip(tied_idx, tied_ref):
            flat[i] = bool(v)
        h, ptr = hashlib.sha256(), 0
        for n in plan.numels:
            h.update(np.packbits(flat[ptr:ptr + n], bitorder="little").tobytes())
            ptr += n
        return h.hexdigest()

    m1, r1 = run_magnitude(plan, info["k"])
    assert float(np.float32(r1["threshold"])) == info["kth_abs"]
    assert r1["n_less"] == info["n_less_round1"] and r1["n_equal"] == info["n_equal_at_kth"]
    assert digest(m1, info["tied_flat_index_round1"], info["tied_ref_mask_round1"]) == info["mask_sha256_round1"]
    zeros, _ = plan.count_zeros(m1)
    assert 100.0 * zeros / plan.total == info["sparsity_after_0.5"]
    m2, r2 = run_magnitude(plan, info["k_round2"], m1)
    assert float(np.float32(r2["threshold"])) == info["kth_abs_round2"]
    assert digest(m2, info["tied_flat_index_round2"], info["tied_ref_mask_round2"]) == info["mask_sha256_round2"]
    zeros, _ = plan.count_zeros(m2)
    assert 100.0 * zeros / plan.total == info["sparsity_after_0.5_then_0.2"]
