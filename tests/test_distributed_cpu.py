"""Multi-rank host logic of the mask build (distributed.ShardedMaskBuilder) on CPU: gloo backend,
world_size 2 and 3, numpy stand-in for the kernels (tests/emul_plan.py), checked against the
oracle.  Also checks the stand-in itself against the oracle on one rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pruning_oracle as PO
from pruning_for_vision_representation_b200 import _lib as L
from pruning_for_vision_representation_b200.distributed import ShardedMaskBuilder, chunk_partition
from tests.emul_plan import NumpyPlan

SIZES = [351, 2808, 25728, 670, 4096, 8192 + 4, 1, 12000]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _weights(seed=0, ties=False):
    rng = np.random.default_rng(seed)
    w = [rng.standard_normal(n).astype(np.float32) for n in SIZES]
    if ties:   # a value repeated in many places on both sides of every rank boundary
        for a in w:
            idx = rng.choice(a.size, size=max(1, a.size // 7), replace=False)
            a[idx] = np.float32(0.25) * rng.choice([-1, 1], size=idx.size).astype(np.float32)
    return w


def _grads(b, seed=100):
    rng = np.random.default_rng(seed + b)
    return [(1e-3 * rng.standard_normal(n)).astype(np.float32) for n in SIZES]


def _flat(arrs):
    return torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrs]).copy())


def _views(flat):
    out, o = [], 0
    for n in SIZES:
        out.append(flat[o:o + n]); o += n
    return out


def _worker(rank, world, port, case, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = NumpyPlan(SIZES)
        builder = ShardedMaskBuilder(plan)
        out = {}
        if case == "snip":
            n_batches = 2 * world
            w = _weights()
            w_flat = _flat(w)
            s_flat = torch.zeros(plan.total)
            plan.bind(L.SLOT_W, _views(w_flat)).bind(L.SLOT_SCORE, _views(s_flat))
            for i, b in enumerate(range(rank * 2, rank * 2 + 2)):
                plan.bind(L.SLOT_G, _views(_flat(_grads(b))))
                plan.score_accumulate(i > 0)
            for sp in (0.9, 1.0, 0.0):
                mask = plan.new_mask()
                s_copy = s_flat.clone()
                plan.bind(L.SLOT_SCORE, _views(s_copy))
                builder.snip_select_emit(s_copy, int(plan.total * sp), mask)
                out[f"mask_{sp}"] = mask.numpy().copy()
                out[f"thr_{sp}"] = plan.result()["threshold"]
            assert n_batches == 2 * world
        elif case == "magnitude":
            w = _weights(ties=True)
            w_flat = _flat(w)
            plan.bind(L.SLOT_W, _views(w_flat))
            old = None
            n_alive = plan.total
            for r, amount in enumerate([0.3, 0.2, 0.5]):
                k = PO.magnitude_k(amount, n_alive)
                new = plan.new_mask()
                builder.magnitude_select_emit(k, old, new)
                out[f"mask_{r}"] = new.numpy().copy()
                out[f"res_{r}"] = {kk: plan.result()[kk] for kk in ("n_less", "n_equal", "quota", "threshold")}
                old = new
                n_alive -= k
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def _run(world, case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return results


def test_chunk_partition():
    assert chunk_partition(10, 3) == [0, 4, 7, 10]
    assert chunk_partition(2, 4) == [0, 1, 2, 2, 2]
    plan = NumpyPlan(SIZES)
    assert plan.chunk_flat_start(0) == 0 and plan.chunk_flat_start(plan.n_chunks) == plan.total
    assert plan.chunk_flat_start(2) == 351 + 2808          # third segment starts at chunk 2
    assert plan.chunk_flat_start(3) == 351 + 2808 + 4096


def test_emulated_plan_matches_oracle_single_rank():
    w = _weights(ties=True)
    plan = NumpyPlan(SIZES)
    plan.bind(L.SLOT_W, _views(_flat(w)))
    old, masks, n_alive = None, None, plan.total
    for amount in (0.3, 0.2, 0.5):
        k = PO.magnitude_k(amount, n_alive)
        new = plan.new_mask()
        plan.select_kth(L.KEY_ABS_W, k, L.MODE_EXACT_K, old)
        plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, new, old)
        masks, info = PO.magnitude_masks(w, masks, amount)
        for got, exp in zip(plan.unpack_mask_host(new), masks):
            assert np.array_equal(got, exp.reshape(-1))
        assert plan.result()["quota"] == info["quota"] and plan.result()["n_equal"] == info["n_equal"]
        old, n_alive = new, n_alive - k


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_snip_matches_oracle(world):
    res = _run(world, "snip")
    w = _weights()
    # expected scores: per-rank sequential accumulation, then ranks added in rank order
    parts = []
    for r in range(world):
        acc = [None] * len(SIZES)
        for b in range(2 * r, 2 * r + 2):
            acc = [PO.snip_score_accumulate(a, wi, gi) for a, wi, gi in zip(acc, w, _grads(b))]
        parts.append(acc)
    scores = parts[0]
    for r in range(1, world):
        scores = [(a + b).astype(np.float32) for a, b in zip(scores, parts[r])]
    flat = np.concatenate(scores)
    plan = NumpyPlan(SIZES)
    for sp in (0.9, 1.0, 0.0):
        thr = PO.snip_threshold(flat, sp)
        exp = PO.snip_masks(scores, thr)
        for r in range(world):
            got = plan.unpack_mask_host(torch.from_numpy(res[r][f"mask_{sp}"]))
            for g, e in zip(got, exp):
                assert np.array_equal(g, e), (sp, r)
            if 0 < int(flat.size * sp) < flat.size:
                assert res[r][f"thr_{sp}"] == thr
        assert all(np.array_equal(res[0][f"mask_{sp}"], res[r][f"mask_{sp}"]) for r in range(world))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_magnitude_ties_match_oracle(world):
    res = _run(world, "magnitude")
    w = _weights(ties=True)
    plan = NumpyPlan(SIZES)
    masks = None
    for r_, amount in enumerate([0.3, 0.2, 0.5]):
        masks, info = PO.magnitude_masks(w, masks, amount)
        for r in range(world):
            got = plan.unpack_mask_host(torch.from_numpy(res[r][f"mask_{r_}"]))
            for g, e in zip(got, masks):
                assert np.array_equal(g, e.reshape(-1)), (r_, r)
            assert res[r][f"res_{r_}"]["quota"] == info["quota"]
            assert res[r][f"res_{r_}"]["n_equal"] == info["n_equal"]
        if r_ == 0:
            assert 0 < info["quota"] < info["n_equal"]          # the cut really falls inside a tied set
