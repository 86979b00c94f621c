// emit_body.cuh — device code of the mask emit shared by emit.cu (k_emit_masks) and select.cu (the finish kernel that
// emits the mask itself once it knows the threshold, b200p_mask_build).
#pragma once
#include "common.cuh"

namespace b200p {

struct EmitArgs {
    const int32_t* chunk_n;
    ChunkTab key_tab;                // |w| or score source
    ChunkTab w_tab;                  // weights (for WEFF output), may equal key_tab
    ChunkTab maskf_tab;              // optional fp32 mask output
    ChunkTab weff_tab;               // optional masked weight output
    const uint32_t* old_mask;        // nullable
    uint32_t* new_mask;
    SelState* st;
    int mode;                        // B200P_MODE_*
    int force;                       // 0 none, 1 keep all, 2 prune all, 3 strict vs forced_threshold
    float forced_threshold;
    int outputs;                     // B200P_EMIT_*
    int vec_ok;
    // emit-by-patch: when the select left a valid provisional mask in `prov`, only the candidates are patched
    int patch; uint32_t* prov; const uint32_t* cand_key; const uint32_t* cand_pos; int64_t n_chunks;
};

// keep decision for one element
__device__ __forceinline__ bool keep_decision(float x, int mode, int force, float thr_f, uint32_t thr_key,
                                              bool ties_pruned) {
    if (force == 1) return true;
    if (force == 2) return false;
    if (mode == B200P_MODE_SNIP_STRICT || force == 3) return x > thr_f;       // NaN -> false (pruned)
    const uint32_t key = key_of(x);
    return key > thr_key || (key == thr_key && !ties_pruned);
}

// ---- K3': emit by patching ---------------------------------------------------------------------------
// The select's sweep already wrote a provisional mask (alive && key >= bracket base) and gathered every
// key inside the bracket with its position.  Once the threshold is known only those candidates can still
// change: clear the bits of the ones that are pruned.  ~2 % of the keys are touched instead of re-reading
// all of them (4.125 B/param -> ~0.3 B/param).  Ties of the chunk where the EXACT_K quota runs out are
// dropped in element order by one warp, exactly like the full emit does.
// The values the patch needs, either read from the select state (stand-alone emit kernel: the select has finished) or
// handed over in registers by the finish kernel that has just derived them (no trip through global memory, no barrier).
struct PatchVals {
    uint32_t cand_count, thr_key, need_ties, tie_resid;
    long long tie_chunk;
    unsigned long long n_kept;
};
__device__ __forceinline__ PatchVals patch_vals_from_state(const SelState* __restrict__ st, int mode) {
    PatchVals v;
    const bool strict = mode == B200P_MODE_SNIP_STRICT;
    v.cand_count = st->cand_count; v.thr_key = st->thr_key; v.need_ties = strict ? 0u : st->need_ties;
    v.tie_resid = st->tie_resid; v.tie_chunk = st->tie_chunk;
    // strict: every tie is pruned, and so are NaN keys although they sort above the threshold (NaN > thr is false, train.py:316;
    // the sweep counted the ones it cleared from the provisional mask into pad_[1])
    v.n_kept = st->n_valid - st->n_less - (strict ? st->n_equal + (st->thr_key != kNanKey ? st->pad_[1] : 0u) : st->quota);
    return v;
}

static __device__ void emit_patch_body(const int32_t* __restrict__ chunk_n, ChunkTab key_tab, const uint32_t* __restrict__ old_mask,
                                SelState* __restrict__ st, const uint32_t* __restrict__ cand_key, const uint32_t* __restrict__ cand_pos,
                                uint32_t* __restrict__ prov, int mode, int64_t n_chunks, const PatchVals& pv,
                                uint32_t* __restrict__ tie_list = nullptr) {
    // tie_list (EXACT_K, few ties): this pass prunes only the keys below the threshold and appends the positions of the tied
    // candidates to the list (tie_list[0] = count); one CTA then prunes the right ones by position (select.cu: tie_list_pick)
    const uint32_t n = pv.cand_count, thr_key = pv.thr_key;
    const bool strict = mode == B200P_MODE_SNIP_STRICT;
    const uint32_t need_ties = (strict || tie_list) ? 0u : pv.need_ties;
    const long long tie_chunk = pv.tie_chunk;
    // eight candidates per thread and iteration: sixteen independent loads in flight, then the (fire-and-forget) bit clears —
    // the loop is bound by load latency, and the fused finish kernel runs it on one CTA per SM
    const uint32_t stride = gridDim.x * kThreads;
    for (uint32_t i0 = blockIdx.x * kThreads + threadIdx.x; i0 < n; i0 += 8 * stride) {
        uint32_t key[8], pos[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t i = i0 + u * stride;
            key[u] = i < n ? __ldg(cand_key + i) : 0xFFFFFFFFu;       // above every threshold: never pruned
            pos[u] = i < n ? __ldg(cand_pos + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long c = pos[u] >> 12;
            bool prune;
            if (strict) prune = key[u] <= thr_key;                    // keep = score > threshold (train.py:316)
            else if (tie_list) {
                prune = key[u] < thr_key;
                if (key[u] == thr_key) {
                    const uint32_t slot = atomicAdd(tie_list, 1u);
                    if (slot < (uint32_t)kTieListCap) tie_list[1 + slot] = pos[u];
                }
            } else {
                const bool ties_pruned = !need_ties || tie_chunk < 0 || c < tie_chunk;
                prune = key[u] < thr_key || (key[u] == thr_key && ties_pruned);
            }
            if (prune) atomicAnd(prov + (size_t)c * kWordsPerChunk + ((pos[u] & 4095u) >> 5), ~(1u << (pos[u] & 31u)));
        }
    }
    if (blockIdx.x != 0) return;
    if (threadIdx.x == 0) st->n_kept = pv.n_kept;
    if (need_ties && tie_chunk >= 0 && tie_chunk < n_chunks && threadIdx.x < 32) {
        // first tie_resid tied + alive keys of this chunk, in element order
        const int lane = threadIdx.x;
        const float* __restrict__ src = chunk_ptr<const float>(key_tab, tie_chunk);
        const int cn = __ldg(chunk_n + tie_chunk);
        const uint32_t* mold = old_mask ? old_mask + tie_chunk * kWordsPerChunk : nullptr;
        uint32_t left = pv.tie_resid;
        for (int wd = 0; wd < kWordsPerChunk && left > 0; ++wd) {
            const int e = wd * 32 + lane;
            bool tie = false;
            if (e < cn) {
                tie = key_of(src[e]) == thr_key;
                if (mold) tie = tie && ((mold[wd] >> lane) & 1u);
            }
            const uint32_t tmask = __ballot_sync(0xFFFFFFFFu, tie);
            if (tmask == 0) continue;
            const uint32_t rank = __popc(tmask & ((1u << lane) - 1u));
            const uint32_t dmask = __ballot_sync(0xFFFFFFFFu, tie && rank < left);
            if (lane == 0) atomicAnd(prov + (size_t)tie_chunk * kWordsPerChunk + wd, ~dmask);
            left -= __popc(dmask);
        }
    }
}

// the full pass: re-reads every key of chunks [c_begin, c_end) (grid-stride over the launching grid)
__device__ __forceinline__ void emit_full_body(const EmitArgs& a, int64_t c_begin, int64_t c_end) {
    __shared__ unsigned long long s_kept;
    if (threadIdx.x == 0) s_kept = 0;
    __syncthreads();
    const int tid = threadIdx.x;
    float thr_f = a.forced_threshold;
    uint32_t thr_key = 0;
    long long tie_chunk = -1;
    uint32_t need_ties = 0, tie_resid = 0;
    if (a.force == 0) {
        thr_f = a.st->threshold; thr_key = a.st->thr_key;
        if (a.mode == B200P_MODE_EXACT_K) { need_ties = a.st->need_ties; tie_chunk = a.st->tie_chunk; tie_resid = a.st->tie_resid; }
    }
    unsigned long long kept = 0;
    const bool want_mf = a.outputs & B200P_EMIT_MASKF, want_wf = a.outputs & B200P_EMIT_WEFF;
    // keep-decision variant, uniform over the launch: 0 strict float compare, 1 integer key compare
    // against a finite threshold (no NaN canonicalisation needed), 2 generic (forced all/none, NaN threshold)
    const int variant = (a.force == 3 || (a.force == 0 && a.mode == B200P_MODE_SNIP_STRICT)) ? 0
                      : (a.force == 0 && thr_key <= 0x7F800000u) ? 1 : 2;

    int64_t c = c_begin + blockIdx.x;
    const float* src = nullptr; int n = 0;
    if (c < c_end) { src = chunk_ptr<const float>(a.key_tab, c); n = __ldg(a.chunk_n + c); }
    for (; c < c_end; ) {
        const int64_t cn = c + gridDim.x;
        const float* srcn = nullptr; int nn = 0;
        if (cn < c_end) { srcn = chunk_ptr<const float>(a.key_tab, cn); nn = __ldg(a.chunk_n + cn); }
        const uint32_t* mold = a.old_mask ? a.old_mask + c * kWordsPerChunk : nullptr;
        uint32_t* mnew = a.new_mask + c * kWordsPerChunk;
        float* mf = want_mf ? chunk_ptr<float>(a.maskf_tab, c) : nullptr;
        float* wf = want_wf ? chunk_ptr<float>(a.weff_tab, c) : nullptr;
        const float* wsrc = want_wf ? chunk_ptr<const float>(a.w_tab, c) : nullptr;
        // ties: pruned everywhere unless the quota runs out at/before this chunk
        const bool ties_pruned = !need_ties || tie_chunk < 0 || c < tie_chunk;

        if (a.vec_ok && n == kChunk) {
            float4 vv[kVecPerThread];
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) vv[j] = ld_nc_f4(src + 4 * (j * kThreads + tid));
            // keep bits of this thread's 16 keys, one variant-specific loop (the variant is launch-uniform)
            uint32_t nibs[kVecPerThread];
            if (variant == 0) {                  // strict float compare (SNIP / forced threshold); NaN -> pruned
#pragma unroll
                for (int j = 0; j < kVecPerThread; ++j)
                    nibs[j] = (vv[j].x > thr_f ? 1u : 0u) | (vv[j].y > thr_f ? 2u : 0u) | (vv[j].z > thr_f ? 4u : 0u) | (vv[j].w > thr_f ? 8u : 0u);
            } else if (variant == 1) {           // integer key compare, finite threshold: raw |x| bits order like the keys
                const int cmp = ties_pruned ? (int)thr_key : (int)thr_key - 1;
#pragma unroll
                for (int j = 0; j < kVecPerThread; ++j)
                    nibs[j] = ((int)(__float_as_uint(vv[j].x) & 0x7FFFFFFFu) > cmp ? 1u : 0u) | ((int)(__float_as_uint(vv[j].y) & 0x7FFFFFFFu) > cmp ? 2u : 0u) |
                              ((int)(__float_as_uint(vv[j].z) & 0x7FFFFFFFu) > cmp ? 4u : 0u) | ((int)(__float_as_uint(vv[j].w) & 0x7FFFFFFFu) > cmp ? 8u : 0u);
            } else {
#pragma unroll
                for (int j = 0; j < kVecPerThread; ++j)
                    nibs[j] = (keep_decision(vv[j].x, a.mode, a.force, thr_f, thr_key, ties_pruned) ? 1u : 0u) |
                              (keep_decision(vv[j].y, a.mode, a.force, thr_f, thr_key, ties_pruned) ? 2u : 0u) |
                              (keep_decision(vv[j].z, a.mode, a.force, thr_f, thr_key, ties_pruned) ? 4u : 0u) |
                              (keep_decision(vv[j].w, a.mode, a.force, thr_f, thr_key, ties_pruned) ? 8u : 0u);
            }
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                const int e = 4 * (j * kThreads + tid);
                const float4 v = vv[j];
                uint32_t oldn = 0xFu;
                if (mold) oldn = nibble_of(__ldg(mold + vec_word_index(j)));
                uint32_t nib = nibs[j];
                nib &= oldn;
                const uint32_t word = gather_nibbles(nib);
                if ((tid & 7) == 0) { mnew[vec_word_index(j)] = word; kept += __popc(word); }
                if (mf) {
                    float4 m; m.x = (nib & 1u) ? 1.f : 0.f; m.y = (nib & 2u) ? 1.f : 0.f;
                    m.z = (nib & 4u) ? 1.f : 0.f; m.w = (nib & 8u) ? 1.f : 0.f;
                    st_f4(mf + e, m);
                }
                if (wf) {
                    const float4 wv = (wsrc == src) ? v : ld_nc_f4(wsrc + e);
                    float4 o; o.x = (nib & 1u) ? wv.x : 0.f; o.y = (nib & 2u) ? wv.y : 0.f;
                    o.z = (nib & 4u) ? wv.z : 0.f; o.w = (nib & 8u) ? wv.w : 0.f;
                    st_f4(wf + e, o);
                }
            }
        } else {
            // scalar path: one 32-element word per warp iteration, lane = bit
            const int warp = tid >> 5, lane = tid & 31;
            for (int wd = warp; wd < kWordsPerChunk; wd += kThreads / 32) {
                const int e = wd * 32 + lane;
                bool keep = false;
                float x = 0.f;
                if (e < n) {
                    x = src[e];
                    keep = keep_decision(x, a.mode, a.force, thr_f, thr_key, ties_pruned);
                    if (mold) keep = keep && ((__ldg(mold + wd) >> lane) & 1u);
                    if (mf) mf[e] = keep ? 1.f : 0.f;
                    if (wf) wf[e] = keep ? wsrc[e] : 0.f;
                }
                const uint32_t word = __ballot_sync(0xFFFFFFFFu, keep);
                if (lane == 0) { mnew[wd] = word; kept += __popc(word); }
            }
        }

        if (need_ties && c == tie_chunk) {
            // The quota runs out inside this chunk: the first tie_resid tied+alive elements (in
            // element order) are pruned, the remaining ties of the chunk stay.  One warp walks
            // the chunk in order; everything above wrote the ties of this chunk as kept.
            __syncthreads();
            if (tid < 32) {
                uint32_t left = tie_resid;
                for (int wd = 0; wd < kWordsPerChunk && left > 0; ++wd) {
                    const int e = wd * 32 + tid;
                    bool tie = false;
                    if (e < n) {
                        tie = key_of(src[e]) == thr_key;
                        if (mold) tie = tie && ((mold[wd] >> tid) & 1u);
                    }
                    const uint32_t tmask = __ballot_sync(0xFFFFFFFFu, tie);
                    if (tmask == 0) continue;
                    const uint32_t rank = __popc(tmask & ((1u << tid) - 1u));
                    const bool drop = tie && rank < left;
                    const uint32_t dmask = __ballot_sync(0xFFFFFFFFu, drop);
                    if (drop) { if (mf) mf[e] = 0.f; if (wf) wf[e] = 0.f; }
                    if (tid == 0) { mnew[wd] &= ~dmask; kept -= __popc(dmask); }
                    left -= __popc(dmask);
                }
            }
            __syncthreads();
        }
        c = cn; src = srcn; n = nn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xFFFFFFFFu, kept, o);
    if ((tid & 31) == 0 && kept) atomicAdd(&s_kept, kept);
    __syncthreads();
    if (tid == 0 && s_kept) atomicAdd(&a.st->n_kept, s_kept);
}

// EmitArgs of a plain (no forced threshold, no fp32 outputs) emit of key_source / mode into new_mask, patching `prov`
inline void fill_patch_emit_args(b200p_plan* p, EmitArgs& a, int key_source, int mode, const uint32_t* d_old_mask, uint32_t* d_new_mask,
                                 uint32_t* prov) {
    const int kslot = key_source == B200P_KEY_ABS_W ? B200P_SLOT_W : B200P_SLOT_SCORE;
    a.patch = 1; a.prov = prov; a.cand_key = p->d_cand_key; a.cand_pos = p->d_cand_pos; a.n_chunks = p->n_chunks;
    a.chunk_n = p->d_chunk_n;
    a.key_tab = p->tab(kslot); a.w_tab = p->tab(B200P_SLOT_W); a.maskf_tab = p->tab(B200P_SLOT_MASKF); a.weff_tab = p->tab(B200P_SLOT_WEFF);
    a.old_mask = d_old_mask; a.new_mask = d_new_mask; a.st = p->d_state;
    a.mode = mode; a.force = 0; a.forced_threshold = 0.f; a.outputs = 0;
    a.vec_ok = p->vec_ok[kslot] ? 1 : 0;
}

}  // namespace b200p
