"""Multi-GPU parity of the peer-memory sharded build, run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank: CUDA IPC windows, ResNet-50-sized set.  (1) magnitude rounds on replicated weights with planted ties,
(2) SNIP with the mini-batches split over the ranks (score kernel writes into the owners' windows), (3) the staged NCCL
fallback.  Each compared bit for bit with the single-GPU build of the same data on every rank.  Prints PASS / FAIL per
item on rank 0, exit code 1 on any failure."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pruning_for_vision_representation_b200 import _lib as L
from pruning_for_vision_representation_b200.distributed import PeerShardedBuilder
from pruning_for_vision_representation_b200.plan import ParamPlan
from pruning_for_vision_representation_b200.shapes import prunable_numels


def views(flat, numels):
    out, off = [], 0
    for n in numels:
        out.append(flat[off:off + n]); off += n
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    numels = prunable_numels(sys.argv[1] if len(sys.argv) > 1 else "resnet50")
    n = sum(numels)
    ok_all = True

    def report(name, ok):
        nonlocal ok_all
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok_all = ok_all and bool(t.item())
        if rank == 0:
            print(f"{'PASS' if t.item() else 'FAIL'}  {name}", flush=True)

    g = torch.Generator(device=dev).manual_seed(1)               # same seed on every rank: replicated weights
    w = torch.randn(n, device=dev, generator=g) * 0.02
    w[::9] = 0.0135; w[4::13] = -0.0135                           # a tied set near the median of |w|
    plan, ref = ParamPlan(numels, dev), ParamPlan(numels, dev)
    plan.bind(L.SLOT_W, views(w, numels)); ref.bind(L.SLOT_W, views(w, numels))
    b = PeerShardedBuilder.from_process_group(plan, score_cap=(plan.n_chunks + world - 1) // world * L.CHUNK)
    old_s = old_r = None
    n_alive = n
    for rnd, amount in enumerate((0.5, 0.2, 0.2)):
        k = round(amount * n_alive)
        m_ref = ref.new_mask()
        ref.mask_build(L.KEY_ABS_W, k, L.MODE_EXACT_K, m_ref, old_r)
        r_ref = ref.result()
        b.magnitude_build(k, old_s)
        res = b.check()
        report(f"magnitude round {rnd} (k={k}, n_equal={r_ref['n_equal']}, quota={r_ref['quota']}, miss={res['miss']})",
               torch.equal(b.mask, m_ref) and res["threshold"] == r_ref["threshold"] and res["quota"] == r_ref["quota"])
        old_r, old_s, n_alive = m_ref, b.mask.clone(), n_alive - k
    # SNIP: 8 mini-batches split over the ranks
    if 8 % world == 0:
        per = 8 // world
        mk = lambda bidx: torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(300 + bidx)) * 1e-3
        mine = [mk(bi) for bi in range(rank * per, (rank + 1) * per)]
        score = torch.zeros(n, device=dev)
        local_tab = plan.pointer_table(L.SLOT_SCORE, views(score, numels))
        gt = [plan.pointer_table(L.SLOT_G, views(x, numels)) for x in mine]
        k = int(n * 0.9)
        b.snip_build(gt, k, score, local_tab)
        res = b.check()
        total, part = torch.empty(n, device=dev), torch.empty(n, device=dev)
        for r in range(world):
            gr = [mk(bi) for bi in range(r * per, (r + 1) * per)]
            ref.bind(L.SLOT_SCORE, views(part if r else total, numels))
            ref.score_accumulate_multi([ref.pointer_table(L.SLOT_G, views(x, numels)) for x in gr])
            if r:
                total.add_(part)
        ref.bind(L.SLOT_SCORE, views(total, numels))
        m_ref = ref.new_mask()
        ref.mask_build(L.KEY_SCORE, k, L.MODE_SNIP_STRICT, m_ref)
        r_ref = ref.result()
        report(f"SNIP scores of the own slice (rank-order sum of {world} parts)", torch.equal(score[b.f0:b.f1], total[b.f0:b.f1]))
        report(f"SNIP mask over {world} GPUs (thr={res['threshold']!r}, kept={res['n_kept']}, miss={res['miss']})",
               torch.equal(b.mask, m_ref) and res["threshold"] == r_ref["threshold"] and res["n_kept"] == r_ref["n_kept"])
    # the staged NCCL fallback on the same data
    k = round(0.37 * n)
    m_ref = ref.new_mask()
    ref.bind(L.SLOT_W, views(w, numels))
    ref.mask_build(L.KEY_ABS_W, k, L.MODE_EXACT_K, m_ref)
    m = plan.new_mask()
    b.fallback.magnitude_select_emit(k, None, m)
    report("staged exact select over NCCL (fallback path)", torch.equal(m, m_ref) and plan.result()["threshold"] == ref.result()["threshold"])
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
