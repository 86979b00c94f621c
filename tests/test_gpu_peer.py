"""The peer-memory sharded mask build (csrc/comm.cuh, b200p_sharded_mask_build) on ONE GPU: G "virtual ranks" are G plans
on the same device whose comm windows are connected by plain pointers (a real run opens them through CUDA IPC; the
kernels do not know the difference).  Each rank runs on its own stream and the stages are interleaved rank by rank, so
the in-kernel all-reduces (last CTA of the sample / sweep kernels), the all-gather of the window histograms and the
mask-word push really wait on each other.  Bar: masks, thresholds and tie bookkeeping bit-identical to the single-plan
build and to the numpy oracle, on every rank."""
import numpy as np
import pytest
import torch

from oracle import pruning_oracle as PO

pytestmark = pytest.mark.gpu

from pruning_for_vision_representation_b200 import _lib as L                       # noqa: E402
from pruning_for_vision_representation_b200.distributed import PeerComm, PeerShardedBuilder  # noqa: E402
from pruning_for_vision_representation_b200.plan import ParamPlan                  # noqa: E402
from pruning_for_vision_representation_b200.shapes import prunable_numels          # noqa: E402

DEV = torch.device("cuda:0")
STAGES = (L.SHARD_SAMPLE, L.SHARD_SWEEP, L.SHARD_FINISH, L.SHARD_TIES, L.SHARD_EMIT, L.SHARD_PUSH)
# the production sequence: sample, sweep, then finish + ties + emit + push as ONE cooperative launch per rank
STAGES_MERGED = (L.SHARD_SAMPLE, L.SHARD_SWEEP, L.SHARD_FINISH | L.SHARD_TIES | L.SHARD_EMIT | L.SHARD_PUSH)


def _views(flat, numels):
    out, off = [], 0
    for n in numels:
        out.append(flat[off:off + n]); off += n
    return out


class VirtualRanks:
    def __init__(self, numels, world, score_cap=0, merged=False):
        self.world = world
        self.stages = STAGES_MERGED if merged else STAGES
        # the ranks' cooperative tail kernels wait on each other: all of them must fit on the one device together
        self.plans = [ParamPlan(numels, DEV).coop_grid_limit(24) for _ in range(world)]
        self.comms = [PeerComm(p, r, world, score_cap) for r, p in enumerate(self.plans)]
        PeerComm.connect_local(self.comms)
        self.builders = [PeerShardedBuilder(p, c) for p, c in zip(self.plans, self.comms)]
        self.streams = [torch.cuda.Stream(device=DEV) for _ in range(world)]

    def each(self, fn):
        torch.cuda.synchronize()
        for r in range(self.world):
            with torch.cuda.stream(self.streams[r]):
                fn(r, self.builders[r])

    def build(self, key_source, old_masks, k, mode):
        """all ranks, stage by stage (the ranks of a real run each issue the whole sequence at once)"""
        torch.cuda.synchronize()
        for stage in self.stages:
            for r in range(self.world):
                with torch.cuda.stream(self.streams[r]):
                    self.builders[r]._build(key_source, None if old_masks is None else old_masks[r], k, mode, stages=stage)
        torch.cuda.synchronize()
        res = [b.check() for b in self.builders]
        return res


def _planted(numels, seed, ties):
    rng = np.random.default_rng(seed)
    w = rng.standard_normal(sum(numels)).astype(np.float32) * 0.05
    if ties:
        w[::7] = np.float32(0.03125)            # a big tied set around the median of |w|
        w[5::11] = np.float32(-0.03125)
    return w


@pytest.mark.parametrize("merged", [False, True])
@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("ties", [False, True])
def test_peer_magnitude_rounds_match_single_plan_and_oracle(world, ties, merged):
    numels = [4096 * 37 + 5, 1000, 4096 * 64, 333, 4096 * 21 + 4095, 77777]
    w = _planted(numels, 3, ties)
    wt = torch.from_numpy(w).to(DEV)
    ref = ParamPlan(numels, DEV)
    ref.bind(L.SLOT_W, _views(wt, numels))
    vr = VirtualRanks(numels, world, merged=merged)
    for p in vr.plans:
        p.bind(L.SLOT_W, _views(wt, numels))
    old_ref, olds, n_alive, omask = None, None, sum(numels), None
    cut_inside_ties = 0
    for amount in (0.5, 0.2, 0.3):
        k = PO.magnitude_k(amount, n_alive)
        new_ref = ref.new_mask()
        ref.mask_build(L.KEY_ABS_W, k, L.MODE_EXACT_K, new_ref, old_ref)
        r_ref = ref.result()
        res = vr.build(L.KEY_ABS_W, olds, k, L.MODE_EXACT_K)
        omask, info = PO.magnitude_masks(_views(w, numels), omask, amount)
        for r in range(world):
            assert res[r]["miss"] == 0
            assert torch.equal(vr.builders[r].mask, new_ref), (world, ties, amount, r)
            for key in ("threshold", "n_less", "n_equal", "quota", "n_valid"):
                assert res[r][key] == r_ref[key], (key, r)
            assert res[r]["n_kept"] == n_alive - k
        assert r_ref["quota"] == info["quota"] and r_ref["n_equal"] == info["n_equal"]
        cut_inside_ties += int(0 < info["quota"] < info["n_equal"])
        for got, exp in zip(ref.unpack_mask_host(new_ref), omask):
            assert np.array_equal(got, exp.reshape(-1))
        old_ref = new_ref
        olds = [b.mask.clone() for b in vr.builders]
        n_alive -= k
    if ties:
        assert cut_inside_ties >= 1          # at least one round had to split a tied set across the ranks' slices


@pytest.mark.parametrize("merged", [True, False])
@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("n_ties", [5, 300])
def test_peer_small_tied_set_across_ranks(world, n_ties, merged):
    """A handful of tied keys spread over the ranks' slices (what real fp32 weight sets have at the cut: ViT-L/16 ties 6
    keys at 50 %): the one-launch sharded tail (merged) resolves them through the tie list, the staged launches through
    the per-chunk table; ties owned by lower ranks count against the quota first either way."""
    numels = [4096 * 37 + 5, 1000, 4096 * 64, 333, 4096 * 21 + 4095, 77777]
    total = sum(numels)
    rng = np.random.default_rng(100 + n_ties)
    w = (rng.standard_normal(total) * 0.05).astype(np.float32)
    tie_val = np.float32(0.0337)
    pos = np.sort(rng.choice(total, n_ties, replace=False))
    w[pos] = tie_val
    n_less = int((np.abs(w) < tie_val).sum())
    wt = torch.from_numpy(w).to(DEV)
    ref = ParamPlan(numels, DEV)
    ref.bind(L.SLOT_W, _views(wt, numels))
    vr = VirtualRanks(numels, world, merged=merged)
    for p in vr.plans:
        p.bind(L.SLOT_W, _views(wt, numels))
    for quota in (1, n_ties // 2, n_ties - 1):
        k = n_less + quota
        new_ref = ref.new_mask()
        ref.mask_build(L.KEY_ABS_W, k, L.MODE_EXACT_K, new_ref, None)
        r_ref = ref.result()
        assert (r_ref["n_less"], r_ref["n_equal"], r_ref["quota"]) == (n_less, n_ties, quota)
        res = vr.build(L.KEY_ABS_W, None, k, L.MODE_EXACT_K)
        exp, _ = PO.magnitude_masks(_views(w, numels), None, int(k))
        for got, e in zip(ref.unpack_mask_host(new_ref), exp):
            assert np.array_equal(got, e.reshape(-1))
        for r in range(world):
            assert res[r]["miss"] == 0 and res[r]["n_kept"] == total - k
            assert torch.equal(vr.builders[r].mask, new_ref), (world, n_ties, quota, r)


@pytest.mark.parametrize("world", [2, 4])
def test_peer_snip_matches_single_plan(world):
    """Batches shard over the ranks; the score kernel writes each chunk's partial scores into the owner's window, the owner
    adds the parts in rank order; then the sharded select.  Reference: the same partial sums added in the same order on one
    plan (identical to the batch-order sum when every rank holds one batch)."""
    numels = [4096 * 50 + 8, 4096 * 3, 100000, 4096 * 40]
    n = sum(numels)
    g = torch.Generator(device=DEV).manual_seed(5)
    wt = torch.randn(n, device=DEV, generator=g) * 0.05
    per = 2
    grads = [[torch.randn(n, device=DEV, generator=g) * 1e-3 for _ in range(per)] for _ in range(world)]
    # one-plan reference: per-rank partials, summed in rank order, then the ordinary mask build
    ref = ParamPlan(numels, DEV)
    parts = torch.empty(world, n, device=DEV)
    ref.bind(L.SLOT_W, _views(wt, numels))
    for r in range(world):
        ref.bind(L.SLOT_SCORE, _views(parts[r], numels))
        ref.score_accumulate_multi([ref.pointer_table(L.SLOT_G, _views(gr, numels)) for gr in grads[r]])
    score_ref = torch.empty(n, device=DEV)
    ref.sum_parts(score_ref, parts.view(-1), world, n, n)
    ref.bind(L.SLOT_SCORE, _views(score_ref, numels))
    k = int(n * 0.9)
    m_ref = ref.new_mask()
    ref.mask_build(L.KEY_SCORE, k, L.MODE_SNIP_STRICT, m_ref)
    r_ref = ref.result()

    vr = VirtualRanks(numels, world, score_cap=n, merged=True)
    scores = [torch.zeros(n, device=DEV) for _ in range(world)]
    local_tabs, gtabs = [], []
    for r, p in enumerate(vr.plans):
        p.bind(L.SLOT_W, _views(wt, numels))
        local_tabs.append(p.pointer_table(L.SLOT_SCORE, _views(scores[r], numels)))
        gtabs.append([p.pointer_table(L.SLOT_G, _views(gr, numels)) for gr in grads[r]])
    # score push of every rank, then the barrier of every rank, then sum + build stage by stage
    def push(r, b):
        p, c = b.plan, b.comm
        import ctypes
        from pruning_for_vision_representation_b200.distributed import _stream
        if b._push_table is None:
            arr = (ctypes.c_int64 * (world + 1))(*b.bounds)
            h = ctypes.c_void_p()
            L.check(p.lib.b200p_comm_score_push_table(c.handle, p.handle, arr, _stream(p), ctypes.byref(h)), "push table")
            b._push_table = b._PtrTable(p, L.SLOT_SCORE, None, [], handle=h)
        p.bind_table(b._push_table)
        rot = b.bounds[(r + 1) % world]
        p.score_accumulate_multi(gtabs[r], accumulate=False, chunk_begin=rot, chunk_end=p.n_chunks)
        if rot > 0:
            p.score_accumulate_multi(gtabs[r], accumulate=False, chunk_begin=0, chunk_end=rot)
    vr.each(push)
    vr.each(lambda r, b: b.comm.barrier())
    def summed(r, b):
        b.plan.sum_parts(scores[r][b.f0:b.f1], b.comm.score_area(), world, b.comm.score_cap, b.f1 - b.f0)
        b.plan.bind_table(local_tabs[r])
    vr.each(summed)
    res = vr.build(L.KEY_SCORE, None, k, L.MODE_SNIP_STRICT)
    for r, b in enumerate(vr.builders):
        assert torch.equal(scores[r][b.f0:b.f1], score_ref[b.f0:b.f1]), r
        assert res[r]["miss"] == 0 and res[r]["threshold"] == r_ref["threshold"] and res[r]["n_kept"] == r_ref["n_kept"]
        assert torch.equal(b.mask, m_ref), r


def test_peer_fullsize_resnet50_magnitude_levels():
    """BASELINE config 5 shape on virtual ranks: ResNet-50-sized set, sparsity 0.5 ... 0.99, 4 ranks."""
    numels = prunable_numels("resnet50")
    n = sum(numels)
    g = torch.Generator(device=DEV).manual_seed(1)
    wt = torch.randn(n, device=DEV, generator=g) * 0.02
    ref = ParamPlan(numels, DEV)
    ref.bind(L.SLOT_W, _views(wt, numels))
    vr = VirtualRanks(numels, 4, merged=True)
    for p in vr.plans:
        p.bind(L.SLOT_W, _views(wt, numels))
    for s in (0.5, 0.8, 0.9, 0.95, 0.99):
        k = round(s * n)
        m = ref.new_mask()
        ref.mask_build(L.KEY_ABS_W, k, L.MODE_EXACT_K, m)
        r_ref = ref.result()
        res = vr.build(L.KEY_ABS_W, None, k, L.MODE_EXACT_K)
        for r, b in enumerate(vr.builders):
            assert res[r]["miss"] == 0 and res[r]["threshold"] == r_ref["threshold"] and torch.equal(b.mask, m), (s, r)


def test_sample_reuse_gives_identical_builds():
    """B200P_OPT_REUSE_SAMPLE: a sparsity sweep over fixed weights derives every bracket after the first from the cached
    sample histogram (one plan and 3 virtual ranks): same masks and thresholds as sampling every time; re-binding the
    weights drops the cache."""
    numels = [4096 * 900 + 3, 4096 * 700, 12345, 4096 * 300]          # 7.8 M keys: large enough for the sampled bracket at every level
    n = sum(numels)
    g = torch.Generator(device=DEV).manual_seed(9)
    wt = torch.randn(n, device=DEV, generator=g) * 0.02
    plain, reuse = ParamPlan(numels, DEV), ParamPlan(numels, DEV).reuse_sample()
    plain.bind(L.SLOT_W, _views(wt, numels)); reuse.bind(L.SLOT_W, _views(wt, numels))
    vr = VirtualRanks(numels, 3)
    for p in vr.plans:
        p.bind(L.SLOT_W, _views(wt, numels)); p.reuse_sample()
    for s in (0.5, 0.9, 0.7, 0.95):
        k = round(s * n)
        m0, m1 = plain.new_mask(), reuse.new_mask()
        plain.mask_build(L.KEY_ABS_W, k, L.MODE_EXACT_K, m0); r0 = plain.result()
        reuse.mask_build(L.KEY_ABS_W, k, L.MODE_EXACT_K, m1); r1 = reuse.result()
        assert torch.equal(m0, m1) and r0["threshold"] == r1["threshold"] and r1["miss"] == r0["miss"] == 0, (s, r0, r1)
        res = vr.build(L.KEY_ABS_W, None, k, L.MODE_EXACT_K)
        for r, b in enumerate(vr.builders):
            assert res[r]["miss"] == 0 and torch.equal(b.mask, m0), (s, r)
    w2 = wt * 1.5                                                   # other weights behind the slot: the cache must not survive the bind
    reuse.bind(L.SLOT_W, _views(w2, numels)); plain.bind(L.SLOT_W, _views(w2, numels))
    k = round(0.7 * n)
    m0, m1 = plain.new_mask(), reuse.new_mask()
    plain.mask_build(L.KEY_ABS_W, k, L.MODE_EXACT_K, m0); reuse.mask_build(L.KEY_ABS_W, k, L.MODE_EXACT_K, m1)
    assert torch.equal(m0, m1) and plain.result()["threshold"] == reuse.result()["threshold"]
