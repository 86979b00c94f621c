"""ParamPlan — Python handle on a b200p_plan (segment tables + select workspace).

PyTorch is plumbing here: it owns the device memory and the stream; every compute call goes
through the C-ABI of include/b200prune.h.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import (CHUNK, WORDS_PER_CHUNK, SLOT_W, SLOT_G, SLOT_SCORE, SLOT_BUF, SLOT_WEFF,
                   SLOT_MASKF, SLOT_WEFF16, SLOT_EMA, MODE_SNIP_STRICT, MODE_EXACT_K, KEY_ABS_W, KEY_SCORE,
                   EMIT_MASKF, EMIT_WEFF, B200PruneError, SelectResult, check)

_SLOT_DTYPE = {SLOT_W: torch.float32, SLOT_G: torch.float32, SLOT_SCORE: torch.float32,
               SLOT_BUF: torch.float32, SLOT_WEFF: torch.float32, SLOT_MASKF: torch.float32,
               SLOT_WEFF16: torch.bfloat16, SLOT_EMA: torch.float32}


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class ParamPlan:
    """Chunk tables for a list of prunable tensors (one segment each, named_modules() order)."""

    def __init__(self, numels, device, cand_capacity=0):
        self.lib = _lib.require_cuda()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise B200PruneError("ParamPlan needs a CUDA device (no CPU fallback)")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.numels = [int(n) for n in numels]
        arr = (ctypes.c_int64 * len(self.numels))(*self.numels)
        handle = ctypes.c_void_p()
        check(self.lib.b200p_plan_create(self.index, len(self.numels), arr, int(cand_capacity),
                                         ctypes.byref(handle)), "plan_create")
        self.handle = handle
        self.total = int(self.lib.b200p_plan_total(handle))
        self.n_chunks = int(self.lib.b200p_plan_num_chunks(handle))
        self.mask_words = int(self.lib.b200p_plan_mask_words(handle))
        self.seg_chunk_start = [int(self.lib.b200p_plan_seg_chunk_start(handle, t))
                                for t in range(len(self.numels) + 1)]
        self.seg_flat_start = [0]
        for n in self.numels:
            self.seg_flat_start.append(self.seg_flat_start[-1] + n)
        self._bound = {}
        self._counts = torch.zeros(2, dtype=torch.int64, device=self.device)

    @classmethod
    def from_tensors(cls, tensors, cand_capacity=0):
        tensors = list(tensors)
        plan = cls([t.numel() for t in tensors], tensors[0].device, cand_capacity)
        return plan

    def set_select_impl(self, impl):
        """'sampled' (default: 1/16 sample -> bracket -> one sweep, exact fallback on a miss) or 'exact'
        (3-pass radix select).  Both return identical results."""
        value = {"sampled": _lib.SELECT_SAMPLED, "exact": _lib.SELECT_EXACT}[impl]
        check(self.lib.b200p_plan_set_option(self.handle, _lib.OPT_SELECT_IMPL, value), "plan_set_option")
        return self

    def coop_grid_limit(self, n_ctas):
        """Cap the grid of this plan's cooperative launches (plans that share one device and wait on each other)."""
        check(self.lib.b200p_plan_set_option(self.handle, _lib.OPT_COOP_GRID, int(n_ctas)), "plan_set_option")
        return self

    def reuse_sample(self, on=True):
        """Selects over unchanged keys (a sparsity sweep over fixed weights) reuse the first one's sample histogram."""
        check(self.lib.b200p_plan_set_option(self.handle, _lib.OPT_REUSE_SAMPLE, 1 if on else 0), "plan_set_option")
        return self

    def close(self):
        if getattr(self, "handle", None):
            self.lib.b200p_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- binding ---------------------------------------------------------------------
    def _validate(self, slot, tensors):
        tensors = list(tensors)
        if len(tensors) != len(self.numels):
            raise B200PruneError(f"bind: expected {len(self.numels)} tensors, got {len(tensors)}")
        want = _SLOT_DTYPE[slot]
        for t, n in zip(tensors, self.numels):
            if not t.is_cuda or t.device.index != self.index:
                raise B200PruneError("bind: tensor is not on the plan's CUDA device (no CPU fallback)")
            if t.dtype != want or not t.is_contiguous() or t.numel() != n:
                raise B200PruneError(f"bind: need contiguous {want} tensors with the plan's element counts")
        return tensors, (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    def pointer_table(self, slot, tensors):
        """Device-resident pointer table for `bind_table`: built once, re-bound with a host-side swap
        (no launch) — one per gradient set in a loop over mini-batches."""
        tensors, ptrs = self._validate(slot, tensors)
        return PtrTable(self, slot, ptrs, tensors)

    def update_tables(self, tables, tensor_lists):
        """Re-point existing pointer tables at new tensor sets (the next build's gradients) in one launch."""
        tables = list(tables)
        slot = tables[0].slot
        arrs = []
        for tab, tensors in zip(tables, tensor_lists):
            if tab.plan is not self or tab.slot != slot:
                raise B200PruneError("update_tables: tables must belong to this plan and one slot")
            tensors, ptrs = self._validate(slot, tensors)
            tab.tensors = tensors
            arrs.append(ptrs)
        tab_arr = (ctypes.c_void_p * len(tables))(*[t.handle for t in tables])
        ptr_arr = (ctypes.c_void_p * len(tables))(*[ctypes.cast(a, ctypes.c_void_p) for a in arrs])
        check(self.lib.b200p_ptrtables_update(tab_arr, ptr_arr, len(tables), slot, _stream_ptr(self.device)), "ptrtables_update")
        self._update_keepalive = arrs
        return self

    def bind_table(self, table):
        if table.plan is not self:
            raise B200PruneError("bind_table: table belongs to another plan")
        check(self.lib.b200p_plan_bind_table(self.handle, table.slot, table.handle), "plan_bind_table")
        self._bound[table.slot] = table   # keeps the table and its tensors alive
        return self

    def bind(self, slot, tensors):
        tensors, ptrs = self._validate(slot, tensors)
        check(self.lib.b200p_plan_bind(self.handle, slot, ptrs, _stream_ptr(self.device)), "plan_bind")
        self._bound[slot] = tensors       # keep the storage alive
        return self

    def bound(self, slot):
        b = self._bound.get(slot)
        return b.tensors if isinstance(b, PtrTable) else b

    def new_mask(self, fill_ones=False):
        """Packed mask tensor (int32 words, chunk-major layout of include/b200prune.h)."""
        m = torch.zeros(self.mask_words, dtype=torch.int32, device=self.device)
        if fill_ones:
            self.emit_masks(KEY_ABS_W if SLOT_W in self._bound else KEY_SCORE, MODE_EXACT_K, m, force=1)
        return m

    # ---- kernels ---------------------------------------------------------------------
    def score_accumulate(self, accumulate, chunk_begin=0, chunk_end=-1):
        check(self.lib.b200p_score_accumulate(self.handle, 1 if accumulate else 0, chunk_begin, chunk_end,
                                              _stream_ptr(self.device)), "score_accumulate")

    def score_accumulate_multi(self, tables, accumulate=False, chunk_begin=0, chunk_end=-1):
        """SCORE (=|+=) sum_b |W * G_b| over the gradient sets in `tables` (PtrTables of slot G) in one pass,
        added in list order: bit-identical to len(tables) score_accumulate calls."""
        tables = list(tables)
        for t in tables:
            if t.plan is not self or t.slot != SLOT_G:
                raise B200PruneError("score_accumulate_multi: need pointer tables of this plan's G slot")
        arr = (ctypes.c_void_p * len(tables))(*[t.handle for t in tables])
        check(self.lib.b200p_score_accumulate_multi(self.handle, arr, len(tables), 1 if accumulate else 0, chunk_begin,
                                                    chunk_end, _stream_ptr(self.device)), "score_accumulate_multi")
        self._multi_keepalive = tables

    def time_sweep(self, on=True):
        """Record CUDA events around every fused score+sweep kernel (measurement only, see kernel_time_ms)."""
        check(self.lib.b200p_plan_set_option(self.handle, _lib.OPT_TIME_SWEEP, 1 if on else 0), "plan_set_option")
        return self

    def kernel_time_ms(self):
        """(mean ms, launches) of the fused score+sweep kernels timed since the last call (synchronises)."""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        check(self.lib.b200p_plan_kernel_time_ms(self.handle, ctypes.byref(ms), ctypes.byref(n)), "plan_kernel_time_ms")
        return float(ms.value), int(n.value)

    def _table_array(self, tables, who):
        tables = list(tables)
        for t in tables:
            if t.plan is not self or t.slot != SLOT_G:
                raise B200PruneError(f"{who}: need pointer tables of this plan's G slot")
        self._multi_keepalive = tables
        return (ctypes.c_void_p * len(tables))(*[t.handle for t in tables]), len(tables)

    def snip_mask_build(self, tables, k, new_mask):
        """SCORE = sum_b |W * G_b|, threshold = k-th smallest, new_mask = score > threshold, in one fused sequence:
        the pass that writes the scores also classifies them (b200p_snip_mask_build)."""
        arr, n = self._table_array(tables, "snip_mask_build")
        check(self.lib.b200p_snip_mask_build(self.handle, arr, n, int(k), _ptr(new_mask), _stream_ptr(self.device)),
              "snip_mask_build")

    def snip_mask_build_refresh(self, tables, tensor_lists, k, new_mask):
        """snip_mask_build for gradient tensors that are new since the tables were filled (fresh backward passes): the
        sample kernel re-points the tables itself — one launch less than update_tables + snip_mask_build."""
        arr, n = self._table_array(tables, "snip_mask_build_refresh")
        arrs = []
        for tab, tensors in zip(tables, tensor_lists):
            tensors, ptrs = self._validate(SLOT_G, tensors)
            tab.tensors = tensors
            arrs.append(ptrs)
        ptr_arr = (ctypes.c_void_p * n)(*[ctypes.cast(a, ctypes.c_void_p) for a in arrs])
        self._update_keepalive = arrs
        check(self.lib.b200p_snip_mask_build_refresh(self.handle, arr, ptr_arr, n, int(k), _ptr(new_mask), _stream_ptr(self.device)),
              "snip_mask_build_refresh")

    def snip_score_select(self, tables, k, prov_target=None):
        """The score + select half of snip_mask_build; follow with emit_masks."""
        arr, n = self._table_array(tables, "snip_score_select")
        check(self.lib.b200p_snip_score_select(self.handle, arr, n, int(k), _ptr(prov_target), _stream_ptr(self.device)),
              "snip_score_select")

    def select_kth(self, key_source, k, mode, old_mask=None):
        check(self.lib.b200p_select_kth(self.handle, key_source, _ptr(old_mask), int(k), mode,
                                        _stream_ptr(self.device)), "select_kth")

    def select_begin(self, k, mode, allow_collect=True):
        check(self.lib.b200p_select_begin(self.handle, int(k), mode, 1 if allow_collect else 0,
                                          _stream_ptr(self.device)), "select_begin")

    def select_hist(self, pass_, key_source, old_mask=None, chunk_begin=0, chunk_end=-1):
        check(self.lib.b200p_select_hist(self.handle, pass_, key_source, _ptr(old_mask), chunk_begin, chunk_end,
                                         _stream_ptr(self.device)), "select_hist")

    def select_scan(self, pass_):
        check(self.lib.b200p_select_scan(self.handle, pass_, _stream_ptr(self.device)), "select_scan")

    def select_ties(self, key_source, old_mask=None, chunk_begin=0, chunk_end=-1, tie_offset=0):
        check(self.lib.b200p_select_ties(self.handle, key_source, _ptr(old_mask), chunk_begin, chunk_end,
                                         int(tie_offset), _stream_ptr(self.device)), "select_ties")

    def select_ties_count(self, key_source, old_mask, chunk_begin, chunk_end, out_count):
        """Ties inside [chunk_begin, chunk_end) -> out_count (int64 device tensor, 1 element)."""
        check(self.lib.b200p_select_ties_count(self.handle, key_source, _ptr(old_mask), chunk_begin, chunk_end,
                                               _ptr(out_count), _stream_ptr(self.device)), "select_ties_count")

    def select_ties_scan(self, chunk_begin, chunk_end, counts, n_before):
        check(self.lib.b200p_select_ties_scan(self.handle, chunk_begin, chunk_end, _ptr(counts), int(n_before),
                                              _stream_ptr(self.device)), "select_ties_scan")

    def sum_parts(self, dst, src, n_parts, part_stride, n):
        """dst[i] = ((src[i] + src[stride+i]) + ...): fixed-order sum of the ranks' score slices."""
        for t in (dst, src):
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise B200PruneError("sum_parts: need contiguous fp32 CUDA tensors (no CPU fallback)")
        if dst.numel() < n or src.numel() < (n_parts - 1) * part_stride + n:
            raise B200PruneError("sum_parts: buffers are too small")
        check(self.lib.b200p_sum_parts(self.index, _ptr(dst), _ptr(src), int(n_parts), int(part_stride), int(n),
                                       _stream_ptr(self.device)), "sum_parts")

    def chunk_flat_start(self, chunk):
        return int(self.lib.b200p_plan_chunk_flat_start(self.handle, int(chunk)))

    def result(self):
        """Select/emit result block (synchronises the current stream)."""
        res = SelectResult()
        check(self.lib.b200p_select_result(self.handle, ctypes.byref(res), _stream_ptr(self.device)), "select_result")
        return res.as_dict()

    def emit_masks(self, key_source, mode, new_mask, old_mask=None, force=0, forced_threshold=0.0,
                   outputs=0, chunk_begin=0, chunk_end=-1):
        check(self.lib.b200p_emit_masks(self.handle, key_source, mode, force, float(forced_threshold),
                                        _ptr(old_mask), _ptr(new_mask), outputs, chunk_begin, chunk_end,
                                        _stream_ptr(self.device)), "emit_masks")

    def mask_build(self, key_source, k, mode, new_mask, old_mask=None):
        """select_kth + emit_masks in one call (the sweep writes the provisional mask straight into new_mask)."""
        check(self.lib.b200p_mask_build(self.handle, key_source, _ptr(old_mask), int(k), mode, _ptr(new_mask),
                                        _stream_ptr(self.device)), "mask_build")

    def count_zeros(self, mask=None, use_weights=True):
        """(zeros of the effective weight, kept bits) as Python ints (one host sync)."""
        check(self.lib.b200p_count_zeros(self.handle, _ptr(mask), _ptr(self._counts), 1 if use_weights else 0,
                                         _stream_ptr(self.device)), "count_zeros")
        z, b = self._counts.tolist()
        return int(z), int(b)

    def mask_pack_from_f32(self, mask):
        check(self.lib.b200p_mask_pack_from_f32(self.handle, _ptr(mask), _stream_ptr(self.device)), "mask_pack")

    def mask_unpack_to_f32(self, mask):
        check(self.lib.b200p_mask_unpack_to_f32(self.handle, _ptr(mask), _stream_ptr(self.device)), "mask_unpack")

    def apply_mask(self, mask, outputs=EMIT_WEFF):
        check(self.lib.b200p_apply_mask(self.handle, _ptr(mask), outputs, _stream_ptr(self.device)), "apply_mask")

    def mask_grads(self, mask):
        check(self.lib.b200p_mask_grads(self.handle, _ptr(mask), _stream_ptr(self.device)), "mask_grads")

    def masked_sgd_step(self, mask, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, flags=0, ctl=None):
        """ctl: optional fp32 CUDA tensor [2] = (gradient multiplier, skip flag), read on the device."""
        check(self.lib.b200p_masked_sgd_step_ctl(self.handle, _ptr(mask), lr, momentum, dampening, weight_decay, flags,
                                                 _ptr(ctl), _stream_ptr(self.device)), "masked_sgd_step")

    def grad_stats(self, mask, out):
        """out (float64 CUDA tensor [2]) = (sum of kept g^2, count of non-finite gradient entries)."""
        check(self.lib.b200p_grad_stats(self.handle, _ptr(mask), _ptr(out), _stream_ptr(self.device)), "grad_stats")

    def ema_update(self, decay, copy=False):
        check(self.lib.b200p_ema_update(self.handle, float(decay), 1 if copy else 0, _stream_ptr(self.device)), "ema_update")

    # ---- workspace views (for collectives between select stages) -----------------------
    def device_view(self, ptr, count, dtype, owner=None):
        """torch view of `count` elements of device memory owned by the library (plan workspace, comm window)."""
        return _device_view(ptr, count, dtype, self.device, owner if owner is not None else self)

    def hist_tensor(self):
        """int64 view [4096] of the plan's global histogram (device memory owned by the plan)."""
        return _device_view(self.lib.b200p_plan_hist_ptr(self.handle), 4096, torch.int64, self.device, self)

    # ---- host-buffer entry points --------------------------------------------------------
    def snip_mask_build_host(self, w_host, g_hosts, k, mask_out_host):
        ptrs = (ctypes.c_void_p * len(g_hosts))(*[g.data_ptr() for g in g_hosts])
        res = SelectResult()
        check(self.lib.b200p_snip_mask_build_host(self.handle, _ptr(w_host), ptrs, len(g_hosts), int(k),
                                                  _ptr(mask_out_host), ctypes.byref(res)), "snip_mask_build_host")
        return res.as_dict()

    def magnitude_mask_build_host(self, w_host, k, mask_out_host, old_mask_host=None):
        res = SelectResult()
        check(self.lib.b200p_magnitude_mask_build_host(self.handle, _ptr(w_host), _ptr(old_mask_host), int(k),
                                                       _ptr(mask_out_host), ctypes.byref(res)),
              "magnitude_mask_build_host")
        return res.as_dict()

    # ---- packed-mask layout helpers (host side, tests / checkpoints) ----------------------
    def unpack_mask_host(self, mask):
        """Packed mask (tensor or ndarray of words) -> list of per-segment bool ndarrays."""
        return unpack_mask_words(np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask), self.numels)


class PtrTable:
    """Handle on a b200p_ptrtable (per-chunk device pointer table of one tensor set)."""

    def __init__(self, plan, slot, ptrs, tensors, handle=None):
        self.plan, self.slot, self.tensors = plan, slot, tensors
        if handle is None:
            handle = ctypes.c_void_p()
            check(plan.lib.b200p_ptrtable_create(plan.handle, slot, ptrs, _stream_ptr(plan.device), ctypes.byref(handle)),
                  "ptrtable_create")
        self.handle = handle

    def close(self):
        if getattr(self, "handle", None) and getattr(self.plan, "handle", None):
            self.plan.lib.b200p_ptrtable_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unpack_mask_words(words, numels):
    words = np.ascontiguousarray(words).view(np.uint32)
    bits = np.unpackbits(words.view(np.uint8), bitorder="little")
    out, chunk = [], 0
    for n in numels:
        nchunks = (n + CHUNK - 1) // CHUNK
        seg = bits[chunk * CHUNK:(chunk + nchunks) * CHUNK]
        out.append(seg[:n].astype(bool))
        chunk += nchunks
    return out


def pack_mask_words(masks, numels):
    """list of per-segment 0/1 arrays -> packed uint32 words (chunk-major, padding bits 0)."""
    total_chunks = sum((n + CHUNK - 1) // CHUNK for n in numels)
    bits = np.zeros(total_chunks * CHUNK, dtype=np.uint8)
    chunk = 0
    for m, n in zip(masks, numels):
        bits[chunk * CHUNK:chunk * CHUNK + n] = np.asarray(m).reshape(-1).astype(bool)
        chunk += (n + CHUNK - 1) // CHUNK
    return np.packbits(bits, bitorder="little").view(np.uint32)


class _DevMem:
    """__cuda_array_interface__ wrapper so torch can view plan-owned device memory."""

    def __init__(self, ptr, nbytes, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _device_view(ptr, count, dtype, device, owner):
    itemsize = torch.empty(0, dtype=dtype).element_size()
    raw = torch.as_tensor(_DevMem(ptr, count * itemsize, owner), device=device)
    return raw.view(dtype)
