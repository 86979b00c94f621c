"""Multi-GPU mask build: one process per GPU, torch.distributed for the plumbing.

    SNIP scores      each rank accumulates |w * g_b| over ITS mini-batches; one exchange
                     (all-to-all of parameter slices, then a fixed rank-order sum on the owner)
                     leaves rank r with the globally summed scores of its own chunk range.
                     Summing in rank order (= mini-batch order) makes the result independent of
                     the collective's reduction topology.
    global threshold parameter-sharded radix select: every rank histograms only its chunk range,
                     the 4096-bin histogram is all-reduced once per radix pass, and every rank
                     runs the identical scan, so all ranks hold the same threshold / tie quota.
    ties (EXACT_K)   per-rank tie counts are all-gathered; rank r skips the ties owned by lower
                     ranks, which keeps the global "lowest flat index first" policy.
    mask emit        each rank emits the packed words of its chunk range; one all-reduce(sum) of
                     the zero-initialised word array gives every replica the full mask.

The reference has no counterpart: under DDP each rank runs `snip_pruning` on its own first batch
and the ranks end up with DIFFERENT masks (train.py:622-628).  Here every rank
finishes with the same mask by construction.

`ShardedMaskBuilder` only talks to a plan-like object (the ParamPlan methods used below), so the
host logic is exercised on CPU with a numpy stand-in for the kernels and the gloo backend
(tests/test_distributed_cpu.py).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib as L
from ._lib import B200PruneError, check


def chunk_partition(n_chunks, world):
    """Contiguous chunk ranges, as even as possible: rank r owns [bounds[r], bounds[r+1])."""
    base, rem = divmod(int(n_chunks), int(world))
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


class ShardedMaskBuilder:
    def __init__(self, plan, group=None):
        self.plan = plan
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.bounds = chunk_partition(plan.n_chunks, self.world)
        self.c0, self.c1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.flat_bounds = [plan.chunk_flat_start(c) for c in self.bounds]
        self.f0, self.f1 = self.flat_bounds[self.rank], self.flat_bounds[self.rank + 1]
        self.split_sizes = [b - a for a, b in zip(self.flat_bounds[:-1], self.flat_bounds[1:])]
        dev = plan.device
        self._recv = torch.empty(self.world * (self.f1 - self.f0), dtype=torch.float32, device=dev)
        self._tie_local = torch.zeros(1, dtype=torch.int64, device=dev)
        self._tie_all = torch.zeros(self.world, dtype=torch.int64, device=dev)
        self._hist = plan.hist_tensor()

    # ---- score exchange ---------------------------------------------------------------------
    def exchange_scores(self, score_flat):
        """score_flat: [N] fp32, this rank's partial scores (segment-concatenated).  Afterwards
        score_flat[f0:f1] holds the sum over ranks, added in rank order.  Returns kernel launches."""
        if self.world == 1:
            return 0
        n_own = self.f1 - self.f0
        dist.all_to_all_single(self._recv, score_flat, output_split_sizes=[n_own] * self.world,
                               input_split_sizes=self.split_sizes, group=self.group)
        self.plan.sum_parts(score_flat[self.f0:self.f1], self._recv, self.world, n_own, n_own)
        return 1

    # ---- sharded select -----------------------------------------------------------------------
    def select(self, key_source, k, mode, old_mask=None):
        """Global k-th smallest over all ranks' chunk ranges.  Returns kernel launches."""
        p = self.plan
        launches = 1
        p.select_begin(k, mode, allow_collect=True)
        for pass_ in range(3):
            p.select_hist(pass_, key_source, old_mask, self.c0, self.c1)
            if self.world > 1:
                dist.all_reduce(self._hist, group=self.group)
            p.select_scan(pass_)
            launches += 2
        if mode == L.MODE_EXACT_K:
            p.select_ties_count(key_source, old_mask, self.c0, self.c1, self._tie_local)
            if self.world > 1:
                dist.all_gather_into_tensor(self._tie_all, self._tie_local, group=self.group)
            else:
                self._tie_all.copy_(self._tie_local)
            p.select_ties_scan(self.c0, self.c1, self._tie_all, self.rank)
            launches += 3
        return launches

    # ---- emit ---------------------------------------------------------------------------------
    def emit(self, key_source, mode, new_mask, old_mask=None, force=0, forced_threshold=0.0):
        """Packed mask of the whole parameter set on every rank.  Returns kernel launches."""
        if self.world > 1:
            new_mask.zero_()
        self.plan.emit_masks(key_source, mode, new_mask, old_mask, force=force, forced_threshold=forced_threshold,
                             chunk_begin=self.c0, chunk_end=self.c1)
        if self.world > 1:
            dist.all_reduce(new_mask, group=self.group)         # disjoint word ranges: sum == or
        return 2 if self.world > 1 else 1

    # ---- whole builds ---------------------------------------------------------------------------
    def snip_select_emit(self, score_flat, k, new_mask):
        """After the local score accumulation: exchange, threshold (train.py:299-307), strict emit
        (train.py:316).  k = int(N * target_sparsity) computed by the caller."""
        n = self.plan.total
        launches = self.exchange_scores(score_flat)
        if k >= n:
            self.plan.select_begin(0, L.MODE_SNIP_STRICT)
            return launches + 1 + self.emit(L.KEY_SCORE, L.MODE_SNIP_STRICT, new_mask, force=3, forced_threshold=float("inf"))
        if k <= 0:
            self.plan.select_begin(0, L.MODE_SNIP_STRICT)
            return launches + 1 + self.emit(L.KEY_SCORE, L.MODE_SNIP_STRICT, new_mask, force=3, forced_threshold=-1.0)
        launches += self.select(L.KEY_SCORE, k, L.MODE_SNIP_STRICT)
        launches += self.emit(L.KEY_SCORE, L.MODE_SNIP_STRICT, new_mask)
        return launches

    def magnitude_select_emit(self, k, old_mask, new_mask):
        """Global magnitude pruning of k alive entries on replicated weights (prune.py:520-540):
        no data exchange, only histograms / tie counts / mask words cross the ranks."""
        if k == 0:
            self.plan.select_begin(0, L.MODE_EXACT_K)
            return 1 + self.emit(L.KEY_ABS_W, L.MODE_EXACT_K, new_mask, old_mask, force=1)
        launches = self.select(L.KEY_ABS_W, k, L.MODE_EXACT_K, old_mask)
        return launches + self.emit(L.KEY_ABS_W, L.MODE_EXACT_K, new_mask, old_mask)

    # ---- host-buffer path (bench.py e2e at N > 1) ------------------------------------------------
    def e2e_host_steps(self, w_host, g_hosts, k, mask_host, steps):
        """Each step: H2D of the weights and of this rank's gradient sets from pinned host memory,
        local accumulate, exchange, select, emit, D2H of the packed mask.  Returns ms per step
        (max over ranks, CUDA events)."""
        p = self.plan
        dev = p.device
        n = p.total
        w_dev = torch.empty(n, dtype=torch.float32, device=dev)
        g_dev = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
        s_dev = torch.empty(n, dtype=torch.float32, device=dev)
        mask = p.new_mask()
        views = lambda flat: [flat[a:b] for a, b in zip(p.seg_flat_start[:-1], p.seg_flat_start[1:])]
        p.bind(L.SLOT_W, views(w_dev)).bind(L.SLOT_SCORE, views(s_dev))
        tabs = [p.pointer_table(L.SLOT_G, views(g)) for g in g_dev]
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)

        def one():
            copied = []
            copy_stream.wait_stream(main)
            with torch.cuda.stream(copy_stream):
                w_dev.copy_(w_host, non_blocking=True)
                ev_w = torch.cuda.Event(); ev_w.record()
            consumed = [None, None]
            for b, gh in enumerate(g_hosts):
                i = b & 1
                with torch.cuda.stream(copy_stream):
                    if consumed[i] is not None:
                        copy_stream.wait_event(consumed[i])
                    g_dev[i].copy_(gh, non_blocking=True)
                    ev = torch.cuda.Event(); ev.record()
                if b == 0:
                    main.wait_event(ev_w)
                main.wait_event(ev)
                p.bind_table(tabs[i])
                p.score_accumulate(b > 0)
                consumed[i] = torch.cuda.Event(); consumed[i].record()
            self.snip_select_emit(s_dev, k, mask)
            mask_host.copy_(mask, non_blocking=True)
            return copied

        one()
        dist.barrier(group=self.group)
        torch.cuda.synchronize(dev)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            one()
        t1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([t0.elapsed_time(t1) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())


# =============================================================================================
# Peer-memory path: the same build with NO NCCL call inside it (csrc/comm.cuh, csrc/comm.cu, csrc/select.cu)
# =============================================================================================
class PeerComm:
    """One rank's window into the peers' memory (b200p_comm).  The kernels of the sharded build push their histograms,
    mask words and partial scores straight into the peers' windows over NVLink; torch.distributed is only used ONCE, to
    hand the 64-byte CUDA IPC handles around."""

    def __init__(self, plan, rank, world, score_cap=0):
        self.lib = L.require_cuda()
        self.plan, self.rank, self.world = plan, int(rank), int(world)
        handle = ctypes.c_void_p()
        check(self.lib.b200p_comm_create(plan.index, self.rank, self.world, plan.mask_words, int(score_cap), ctypes.byref(handle)),
              "comm_create")
        self.handle = handle
        self.score_cap = int(self.lib.b200p_comm_score_cap(handle))

    @classmethod
    def from_process_group(cls, plan, group=None, score_cap=0):
        """One comm per rank of `group` (one process per GPU, one node), windows opened through CUDA IPC."""
        group = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        comm = cls(plan, rank, world, score_cap)
        if world > 1:
            buf = ctypes.create_string_buffer(64)
            check(comm.lib.b200p_comm_ipc_handle(comm.handle, buf), "comm_ipc_handle")
            mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(plan.device)
            allh = torch.empty(world * 64, dtype=torch.uint8, device=plan.device)
            dist.all_gather_into_tensor(allh, mine, group=group)
            raw = bytes(allh.cpu().numpy().tobytes())
            check(comm.lib.b200p_comm_connect_ipc(comm.handle, raw), "comm_connect_ipc")
            dist.barrier(group=group)           # every rank has mapped every window before the first kernel touches one
        return comm

    @staticmethod
    def connect_local(comms):
        """Comms of plans that live in ONE process (virtual ranks on one device, tests): plain pointers, no IPC."""
        wins = [int(c.lib.b200p_comm_window(c.handle)) for c in comms]
        arr = (ctypes.c_void_p * len(wins))(*wins)
        for c in comms:
            check(c.lib.b200p_comm_connect_local(c.handle, arr), "comm_connect_local")

    def mask_tensor(self):
        """int32 view [mask_words] of the full packed mask inside the own window (valid after a build's all-gather)."""
        return self.plan.device_view(self.lib.b200p_comm_mask_ptr(self.handle), self.plan.mask_words, torch.int32, self)

    def score_area(self):
        return self.plan.device_view(self.lib.b200p_comm_score_ptr(self.handle), self.world * self.score_cap, torch.float32, self)

    def barrier(self):
        check(self.lib.b200p_comm_barrier(self.handle, _stream(self.plan)), "comm_barrier")

    def error(self):
        return int(self.lib.b200p_comm_error(self.handle, _stream(self.plan)))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.b200p_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _stream(plan):
    return ctypes.c_void_p(torch.cuda.current_stream(plan.device).cuda_stream)


class PeerShardedBuilder:
    """Parameter-sharded mask build over peer memory.  Rank r owns the contiguous chunk range bounds[r]:bounds[r+1];
    every rank ends with the full packed mask in `self.mask` (a view of its window), bit-identical to the single-GPU
    build.  Per build: sample -> sweep of the own slice -> exact key -> ties -> patch emit -> mask push, with the three
    small exchanges done by the last CTA of the kernel that produced the data (no launch, no NCCL, no host sync).

    SNIP: `snip_build` first folds this rank's gradient sets into partial scores that the score kernel writes straight
    into the OWNER's window (score pass + all-to-all in one kernel), then the owner adds the parts in rank order."""

    def __init__(self, plan, comm, fallback=None):
        from .plan import PtrTable
        self.plan, self.comm, self.fallback = plan, comm, fallback
        self.world, self.rank = comm.world, comm.rank
        self.bounds = chunk_partition(plan.n_chunks, self.world)
        self.c0, self.c1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.flat_bounds = [plan.chunk_flat_start(c) for c in self.bounds]
        self.f0, self.f1 = self.flat_bounds[self.rank], self.flat_bounds[self.rank + 1]
        self.mask = comm.mask_tensor()
        self._push_table = None
        self._PtrTable = PtrTable
        self._last = None

    @classmethod
    def from_process_group(cls, plan, group=None, score_cap=0):
        """One process per GPU: comm windows opened through CUDA IPC, and the staged exact select over NCCL
        (ShardedMaskBuilder) as the fallback `check()` runs when the sampled bracket missed (small or degenerate key sets)."""
        comm = PeerComm.from_process_group(plan, group, score_cap)
        return cls(plan, comm, fallback=ShardedMaskBuilder(plan, group))

    # ---- magnitude (replicated weights: nothing but histograms, 4 KB of counts and mask words crosses the ranks) -------
    def magnitude_build(self, k, old_mask=None):
        p = self.plan
        if k == 0:                                              # prune.py:533: mask unchanged
            p.select_begin(0, L.MODE_EXACT_K)
            p.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, self.mask, old_mask, force=1, chunk_begin=self.c0, chunk_end=self.c1)
            self._allgather()
            self._last = None
            return 2
        self._build(L.KEY_ABS_W, old_mask, k, L.MODE_EXACT_K)
        return 7

    # ---- SNIP ----------------------------------------------------------------------------------------------------------
    def snip_build(self, g_tables, k, score_flat, local_score_table):
        """g_tables: this rank's gradient sets (PtrTables, mini-batch order); score_flat: [N] fp32 whose slice [f0:f1)
        receives the summed scores; local_score_table: PtrTable of the SCORE slot over score_flat."""
        p, c = self.plan, self.comm
        if c.score_cap < max(b - a for a, b in zip(self.flat_bounds[:-1], self.flat_bounds[1:])):
            raise B200PruneError("snip_build: the comm was created without a score area (score_cap)")
        if self._push_table is None:
            arr = (ctypes.c_int64 * (self.world + 1))(*self.bounds)
            h = ctypes.c_void_p()
            check(p.lib.b200p_comm_score_push_table(c.handle, p.handle, arr, _stream(p), ctypes.byref(h)), "comm_score_push_table")
            self._push_table = self._PtrTable(p, L.SLOT_SCORE, None, [], handle=h)
        p.bind_table(self._push_table)
        # every rank starts with the chunks of its right-hand neighbour, so that at any moment the G ranks store into G
        # different windows (an all-to-all, not G incasts)
        rot = self.bounds[(self.rank + 1) % self.world]
        p.score_accumulate_multi(g_tables, accumulate=False, chunk_begin=rot, chunk_end=p.n_chunks)
        if rot > 0:
            p.score_accumulate_multi(g_tables, accumulate=False, chunk_begin=0, chunk_end=rot)
        c.barrier()                                             # every rank's parts have landed in every window
        n_own = self.f1 - self.f0
        p.sum_parts(score_flat[self.f0:self.f1], c.score_area(), self.world, c.score_cap, n_own)
        p.bind_table(local_score_table)
        n = p.total
        if k >= n or k <= 0:                                    # train.py:300-303
            p.select_begin(0, L.MODE_SNIP_STRICT)
            p.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, self.mask, None, force=3,
                         forced_threshold=float("inf") if k >= n else -1.0, chunk_begin=self.c0, chunk_end=self.c1)
            self._allgather()
            self._last = None
            return 6
        self._build(L.KEY_SCORE, None, k, L.MODE_SNIP_STRICT)
        return 10

    def _allgather(self):
        check(self.plan.lib.b200p_comm_mask_allgather(self.comm.handle, self.plan.handle, self.c0, self.c1, _stream(self.plan)),
              "comm_mask_allgather")

    def _build(self, key_source, old_mask, k, mode, stages=0):
        p = self.plan
        check(p.lib.b200p_sharded_mask_build(p.handle, self.comm.handle, key_source,
                                             ctypes.c_void_p(old_mask.data_ptr()) if old_mask is not None else None,
                                             int(k), mode, self.c0, self.c1, int(stages), _stream(p)), "sharded_mask_build")
        self._last = (key_source, old_mask, int(k), mode)

    def check(self):
        """Result block of the last build (synchronises, like the reference's `.item()` for its threshold print).  Raises
        if a peer never arrived; reruns the staged exact select (NCCL histogram all-reduces) if the bracket missed."""
        err = self.comm.error()
        if err:
            raise B200PruneError(f"sharded build: wait on channel {err - 1} timed out (a peer rank never arrived)")
        res = self.plan.result()
        if res["miss"] and self._last is not None:
            if self.fallback is None:
                raise B200PruneError("sharded build: the sampled bracket missed and no staged fallback builder was given")
            key_source, old_mask, k, mode = self._last
            self.fallback.select(key_source, k, mode, old_mask)
            self.fallback.emit(key_source, mode, self.mask, old_mask)
            res = self.plan.result()
        strict = self._last is not None and self._last[3] == L.MODE_SNIP_STRICT
        res["n_kept"] = res["n_valid"] - res["n_less"] - (res["n_equal"] if strict else res["quota"])
        return res
