"""The staged (multi-GPU) select / tie / emit entry points on one GPU: two "virtual ranks" own the two
halves of the chunk range and share the plan's histogram (what the NCCL all-reduce produces), plus the
real ShardedMaskBuilder on a 1-rank process group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist

from oracle import pruning_oracle as PO

pytestmark = pytest.mark.gpu

from pruning_for_vision_representation_b200 import _lib as L                       # noqa: E402
from pruning_for_vision_representation_b200.distributed import ShardedMaskBuilder, chunk_partition  # noqa: E402
from pruning_for_vision_representation_b200.plan import ParamPlan                  # noqa: E402
from tests.test_distributed_cpu import SIZES, _weights, _grads                     # noqa: E402

DEV = torch.device("cuda:0")


def _dev(arrs):
    return [torch.from_numpy(a.copy()).to(DEV) for a in arrs]


@pytest.mark.parametrize("world", [2, 3])
def test_virtual_ranks_magnitude_with_ties(world):
    """Histogram mode (allow_collect=False): the candidate buffer of collect mode is per rank in a
    real multi-GPU run and cannot be shared by virtual ranks; collect mode is covered by
    test_builder_world1_* and the multi-GPU bench."""
    w = _weights(ties=True)
    plan = ParamPlan(SIZES, DEV)
    plan.bind(L.SLOT_W, _dev(w))
    bounds = chunk_partition(plan.n_chunks, world)
    masks, old, n_alive = None, None, plan.total
    for amount in (0.3, 0.2, 0.5):
        k = PO.magnitude_k(amount, n_alive)
        plan.select_begin(k, L.MODE_EXACT_K, allow_collect=False)
        for p in range(3):
            for r in range(world):
                plan.select_hist(p, L.KEY_ABS_W, old, bounds[r], bounds[r + 1])     # histograms add up = all-reduce
            plan.select_scan(p)
        counts = torch.zeros(world, dtype=torch.int64, device=DEV)
        for r in range(world):                                                      # = all-gather of the tie counts
            plan.select_ties_count(L.KEY_ABS_W, old, bounds[r], bounds[r + 1], counts[r:r + 1])
        new = plan.new_mask()
        for r in range(world):
            plan.select_ties_scan(bounds[r], bounds[r + 1], counts, r)
            plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, new, old, chunk_begin=bounds[r], chunk_end=bounds[r + 1])
        masks, info = PO.magnitude_masks(w, masks, amount)
        res = plan.result()
        assert res["quota"] == info["quota"] and res["n_equal"] == info["n_equal"]
        assert int(counts.sum()) == (info["n_equal"] if info["quota"] < info["n_equal"] else 0)
        for got, exp in zip(plan.unpack_mask_host(new), masks):
            assert np.array_equal(got, exp.reshape(-1))
        old, n_alive = new, n_alive - k


def test_sum_parts_fixed_order():
    rng = np.random.default_rng(3)
    for n, parts, off in ((10007, 3, 0), (4096 * 5, 8, 0), (777, 2, 1)):
        src = rng.standard_normal((parts, n)).astype(np.float32) * np.float32(1e-3)
        exp = src[0].copy()
        for p in range(1, parts):
            exp = (exp + src[p]).astype(np.float32)
        s = torch.from_numpy(src).to(DEV).reshape(-1)
        big = torch.zeros(n + 8, device=DEV)
        dst = big[off:off + n]
        plan = ParamPlan([n], DEV)
        plan.sum_parts(dst, s, parts, n, n)
        assert np.array_equal(dst.cpu().numpy(), exp)


def test_builder_world1_snip_and_magnitude():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=DEV)
    try:
        w = _weights(ties=True)
        plan = ParamPlan(SIZES, DEV)
        wt = _dev(w)
        s_flat = torch.zeros(plan.total, device=DEV)
        st = [s_flat[a:b] for a, b in zip(plan.seg_flat_start[:-1], plan.seg_flat_start[1:])]
        plan.bind(L.SLOT_W, wt).bind(L.SLOT_SCORE, st)
        builder = ShardedMaskBuilder(plan)
        grads = [_grads(b) for b in range(3)]
        for b, g in enumerate(grads):
            plan.bind(L.SLOT_G, _dev(g))
            plan.score_accumulate(b > 0)
        mask = plan.new_mask()
        builder.snip_select_emit(s_flat, int(plan.total * 0.8), mask)
        exp, thr, _ = PO.snip_pruning(w, grads, 0.8)
        assert plan.result()["threshold"] == np.float32(thr)
        for got, e in zip(plan.unpack_mask_host(mask), exp):
            assert np.array_equal(got, e.reshape(-1))
        masks, old, n_alive = None, None, plan.total
        for amount in (0.3, 0.2):
            k = PO.magnitude_k(amount, n_alive)
            new = plan.new_mask()
            builder.magnitude_select_emit(k, old, new)
            masks, info = PO.magnitude_masks(w, masks, amount)
            for got, e in zip(plan.unpack_mask_host(new), masks):
                assert np.array_equal(got, e.reshape(-1))
            old, n_alive = new, n_alive - k
    finally:
        dist.destroy_process_group()
