"""Multi-GPU parity check (run under torchrun, NCCL): the sharded mask build of
distributed.ShardedMaskBuilder against the single-GPU kernels on the same data, bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pruning_for_vision_representation_b200 import _lib as L
from pruning_for_vision_representation_b200.distributed import ShardedMaskBuilder
from pruning_for_vision_representation_b200.plan import ParamPlan
from pruning_for_vision_representation_b200.shapes import prunable_numels


def views(flat, numels):
    out, o = [], 0
    for n in numels:
        out.append(flat[o:o + n]); o += n
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    numels = prunable_numels("resnet18") + [351, 2808, 670, 1]          # plus awkward partial chunks
    n = sum(numels)
    g = torch.Generator(device=dev).manual_seed(1)
    w = torch.randn(n, device=dev, generator=g) * 0.02
    w[torch.randperm(n, device=dev, generator=g)[: n // 9]] = 0.0078125   # planted ties around the 10-20 % quantile
    per_rank = 2
    grads = [torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(300 + b)) * 1e-3
             for b in range(world * per_rank)]
    plan = ParamPlan(numels, dev)
    s = torch.empty(n, device=dev)
    plan.bind(L.SLOT_W, views(w, numels)).bind(L.SLOT_SCORE, views(s, numels))
    builder = ShardedMaskBuilder(plan)
    ok = True
    for sparsity in (0.9, 0.5, 1.0, 0.0):
        k = int(n * sparsity)
        # --- sharded
        for i, b in enumerate(range(rank * per_rank, (rank + 1) * per_rank)):
            plan.bind(L.SLOT_G, views(grads[b], numels)); plan.score_accumulate(i > 0)
        mask_d = plan.new_mask()
        builder.snip_select_emit(s, k, mask_d)
        thr_d = plan.result()["threshold"]
        # --- single GPU, same summation tree: per-rank partials, then rank-order sum
        parts = torch.empty(world, n, device=dev)
        for r in range(world):
            plan.bind(L.SLOT_SCORE, views(parts[r], numels))
            for i, b in enumerate(range(r * per_rank, (r + 1) * per_rank)):
                plan.bind(L.SLOT_G, views(grads[b], numels)); plan.score_accumulate(i > 0)
        plan.sum_parts(s, parts.view(-1), world, n, n)
        plan.bind(L.SLOT_SCORE, views(s, numels))
        mask_s = plan.new_mask()
        if k >= n:
            plan.select_begin(0, L.MODE_SNIP_STRICT); plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, mask_s, force=3, forced_threshold=float("inf"))
        elif k <= 0:
            plan.select_begin(0, L.MODE_SNIP_STRICT); plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, mask_s, force=3, forced_threshold=-1.0)
        else:
            plan.select_kth(L.KEY_SCORE, k, L.MODE_SNIP_STRICT); plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, mask_s)
        thr_s = plan.result()["threshold"]
        same = torch.equal(mask_d, mask_s) and (not (0 < k < n) or thr_d == thr_s)
        ok &= same
        if rank == 0:
            print(f"snip sparsity={sparsity}: thr {thr_d} vs {thr_s}, kept {int(plan.count_zeros(mask_d, use_weights=False)[1])}, identical={same}", flush=True)
    # --- magnitude, iterative, ties
    old_d = old_s = None
    n_alive = n
    for amount in (0.15, 0.2, 0.5):
        k = round(amount * n_alive)
        new_d = plan.new_mask(); builder.magnitude_select_emit(k, old_d, new_d); rd = plan.result()
        new_s = plan.new_mask(); plan.select_kth(L.KEY_ABS_W, k, L.MODE_EXACT_K, old_s); plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, new_s, old_s); rs = plan.result()
        same = torch.equal(new_d, new_s) and rd["quota"] == rs["quota"] and rd["n_equal"] == rs["n_equal"]
        ok &= same
        if rank == 0:
            print(f"magnitude amount={amount}: k={k} n_equal={rd['n_equal']} quota={rd['quota']} identical={same}", flush=True)
        old_d, old_s, n_alive = new_d, new_s, n_alive - k
    # every rank holds the same mask
    ref = old_d.clone(); dist.broadcast(ref, 0)
    ok &= torch.equal(ref, old_d)
    t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
