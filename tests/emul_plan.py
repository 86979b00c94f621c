"""numpy stand-in for ParamPlan — TEST INFRASTRUCTURE.

Restates the integer logic of the CUDA kernels (31-bit keys, 12/12/7-bit digits, staged
histogram select, lowest-index-first tie quota, chunk-major packed masks) on CPU tensors so that
the multi-rank host logic of `distributed.ShardedMaskBuilder` can run under gloo without a GPU.
It has the ParamPlan methods the builder calls and nothing else.
"""
import numpy as np
import torch

from pruning_for_vision_representation_b200 import _lib as L
from pruning_for_vision_representation_b200.plan import pack_mask_words, unpack_mask_words

CHUNK = L.CHUNK
WORDS = L.WORDS_PER_CHUNK
NAN_KEY = 0x7FFFFFFF
SHIFT = (19, 7, 0)
BINS = (4096, 4096, 128)
PMASK = (0, 0x7FF80000, 0x7FFFFF80)


def key_of(x):
    u = np.asarray(x, np.float32).view(np.uint32) & np.uint32(0x7FFFFFFF)
    return np.where(u > 0x7F800000, np.uint32(NAN_KEY), u).astype(np.uint32)


class NumpyPlan:
    def __init__(self, numels):
        self.numels = [int(n) for n in numels]
        self.device = torch.device("cpu")
        self.total = sum(self.numels)
        self.seg_chunk_start = [0]
        self.seg_flat_start = [0]
        for n in self.numels:
            self.seg_chunk_start.append(self.seg_chunk_start[-1] + (n + CHUNK - 1) // CHUNK)
            self.seg_flat_start.append(self.seg_flat_start[-1] + n)
        self.n_chunks = self.seg_chunk_start[-1]
        self.mask_words = self.n_chunks * WORDS
        self._bound = {}
        self._hist = torch.zeros(4096, dtype=torch.int64)
        self.state = {}
        # padded index space: element e of segment t lives at seg_chunk_start[t]*CHUNK + e
        self.valid = np.zeros(self.n_chunks * CHUNK, dtype=bool)
        for t, n in enumerate(self.numels):
            o = self.seg_chunk_start[t] * CHUNK
            self.valid[o:o + n] = True

    # ---- plumbing ----
    def pointer_table(self, slot, tensors):
        tensors = list(tensors)
        assert [t.numel() for t in tensors] == self.numels
        return (slot, None, tensors)

    def bind_table(self, table):
        self._bound[table[0]] = table[2]
        return self

    def bind(self, slot, tensors):
        return self.bind_table(self.pointer_table(slot, tensors))

    def chunk_flat_start(self, c):
        if c == self.n_chunks:
            return self.total
        t = max(i for i in range(len(self.numels)) if self.seg_chunk_start[i] <= c)
        return self.seg_flat_start[t] + (c - self.seg_chunk_start[t]) * CHUNK

    def hist_tensor(self):
        return self._hist

    def new_mask(self, fill_ones=False):
        m = torch.zeros(self.mask_words, dtype=torch.int32)
        if fill_ones:
            m.copy_(torch.from_numpy(pack_mask_words([np.ones(n, bool) for n in self.numels], self.numels).view(np.int32)))
        return m

    def _padded(self, slot):
        out = np.zeros(self.n_chunks * CHUNK, dtype=np.float32)
        for t, ten in enumerate(self._bound[slot]):
            o = self.seg_chunk_start[t] * CHUNK
            out[o:o + self.numels[t]] = ten.detach().numpy().reshape(-1)
        return out

    def _keys(self, key_source):
        return key_of(self._padded(L.SLOT_W if key_source == L.KEY_ABS_W else L.SLOT_SCORE))

    def _alive(self, old_mask):
        if old_mask is None:
            return self.valid.copy()
        bits = np.unpackbits(old_mask.numpy().view(np.uint8), bitorder="little").astype(bool)
        return bits & self.valid

    # ---- kernels ----
    def score_accumulate(self, accumulate, chunk_begin=0, chunk_end=-1):
        for w, g, s in zip(self._bound[L.SLOT_W], self._bound[L.SLOT_G], self._bound[L.SLOT_SCORE]):
            r = (w.detach() * g.detach()).abs()
            s.copy_(s + r if accumulate else r)

    def sum_parts(self, dst, src, n_parts, part_stride, n):
        acc = src[:n].clone()
        for p in range(1, n_parts):
            acc = acc + src[p * part_stride:p * part_stride + n]
        dst[:n].copy_(acc)

    def select_begin(self, k, mode, allow_collect=True):
        self.state = dict(k=int(k), k_request=int(k), mode=mode, prefix=0, n_less=0, n_valid=0, n_equal=0, quota=0,
                          need_ties=0, thr_key=0, threshold=0.0, tie_chunk=-1, tie_resid=0, n_kept=0, invalid=False)
        self._hist.zero_()

    def select_hist(self, pass_, key_source, old_mask=None, chunk_begin=0, chunk_end=-1):
        if chunk_end < 0:
            chunk_end = self.n_chunks
        sl = slice(chunk_begin * CHUNK, chunk_end * CHUNK)
        keys = self._keys(key_source)[sl]
        sel = self._alive(old_mask)[sl] & ((keys & np.uint32(PMASK[pass_])) == np.uint32(self.state["prefix"]))
        digits = (keys[sel] >> np.uint32(SHIFT[pass_])) & np.uint32(BINS[pass_] - 1)
        self._hist[:BINS[pass_]] += torch.from_numpy(np.bincount(digits, minlength=BINS[pass_]).astype(np.int64))

    def select_scan(self, pass_):
        st = self.state
        h = self._hist.numpy()[:BINS[pass_]].copy()
        total = int(h.sum())
        if pass_ == 0:
            st["n_valid"] = total
        k = st["k"]
        if k == 0 or k > total:
            st.update(prefix=NAN_KEY, thr_key=NAN_KEY, threshold=float("nan"), n_equal=0, quota=0, need_ties=0, invalid=True)
        else:
            cum = np.cumsum(h)
            b = int(np.searchsorted(cum, k, side="left"))
            before = int(cum[b] - h[b])
            st["n_less"] += before
            st["k"] = k - before
            st["prefix"] |= b << SHIFT[pass_]
            if pass_ == 2:
                st["thr_key"] = st["prefix"]
                st["threshold"] = float(np.array([st["prefix"] if st["prefix"] != NAN_KEY else 0x7FC00000], np.uint32).view(np.float32)[0])
                st["n_equal"] = int(h[b])
                st["quota"] = k - before
                st["need_ties"] = int(st["mode"] == L.MODE_EXACT_K and st["quota"] < st["n_equal"])
                st["tie_chunk"], st["tie_resid"] = -1, 0
        self._hist.zero_()

    def _tie_flags(self, key_source, old_mask):
        return self._alive(old_mask) & (self._keys(key_source) == np.uint32(self.state["thr_key"]))

    def select_ties_count(self, key_source, old_mask, chunk_begin, chunk_end, out_count):
        self._ties = self._tie_flags(key_source, old_mask) if self.state["need_ties"] else None
        n = 0 if self._ties is None else int(self._ties[chunk_begin * CHUNK:chunk_end * CHUNK].sum())
        out_count.fill_(n)

    def select_ties_scan(self, chunk_begin, chunk_end, counts, n_before):
        st = self.state
        if not st["need_ties"]:
            return
        offset = int(counts[:n_before].sum())
        target = st["quota"] - offset
        per_chunk = self._ties.reshape(self.n_chunks, CHUNK).sum(axis=1)[chunk_begin:chunk_end]
        if target <= 0:
            st["tie_chunk"], st["tie_resid"] = chunk_begin, 0
        elif target > int(per_chunk.sum()):
            st["tie_chunk"], st["tie_resid"] = chunk_end, 0
        else:
            cum = np.cumsum(per_chunk)
            i = int(np.searchsorted(cum, target, side="left"))
            st["tie_chunk"] = chunk_begin + i
            st["tie_resid"] = int(target - (cum[i] - per_chunk[i]))

    def select_kth(self, key_source, k, mode, old_mask=None):
        self.select_begin(k, mode)
        for p in range(3):
            self.select_hist(p, key_source, old_mask)
            self.select_scan(p)
        if mode == L.MODE_EXACT_K:
            cnt = torch.zeros(1, dtype=torch.int64)
            self.select_ties_count(key_source, old_mask, 0, self.n_chunks, cnt)
            self.select_ties_scan(0, self.n_chunks, cnt, 0)

    def emit_masks(self, key_source, mode, new_mask, old_mask=None, force=0, forced_threshold=0.0,
                   outputs=0, chunk_begin=0, chunk_end=-1):
        if chunk_end < 0:
            chunk_end = self.n_chunks
        st = self.state
        slot = L.SLOT_W if key_source == L.KEY_ABS_W else L.SLOT_SCORE
        x = self._padded(slot)
        alive = self._alive(old_mask)
        if force == 1:
            keep = np.ones_like(alive)
        elif force == 2:
            keep = np.zeros_like(alive)
        elif force == 3 or mode == L.MODE_SNIP_STRICT:
            thr = np.float32(forced_threshold if force == 3 else st["threshold"])
            with np.errstate(invalid="ignore"):
                keep = x > thr
        else:
            keys = key_of(x)
            thr = np.uint32(st["thr_key"])
            keep = keys > thr
            tie = (keys == thr) & alive
            if st["need_ties"] and st["tie_chunk"] >= 0:
                chunk_of = np.arange(x.size) // CHUNK
                keep |= tie & (chunk_of > st["tie_chunk"])
                inside = np.flatnonzero(tie & (chunk_of == st["tie_chunk"]))
                keep[inside[st["tie_resid"]:]] = True
        keep &= alive
        bits = np.zeros_like(keep)
        sl = slice(chunk_begin * CHUNK, chunk_end * CHUNK)
        bits[sl] = keep[sl]
        words = np.packbits(bits.astype(np.uint8), bitorder="little").view(np.int32)
        wsl = slice(chunk_begin * WORDS, chunk_end * WORDS)
        new_mask.numpy()[wsl] = words[wsl]
        st["n_kept"] = int(bits.sum())

    def result(self):
        return dict(self.state)

    def unpack_mask_host(self, mask):
        return unpack_mask_words(mask.numpy(), self.numels)
