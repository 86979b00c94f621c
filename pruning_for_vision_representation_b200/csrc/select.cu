// select.cu — K2: global k-th smallest key by multi-pass MSD radix select.
//
// Replaces the full torch.sort of train.py:306-307 (SNIP) and the torch.topk of
// torch/nn/utils/prune.py:536 (global magnitude): only the k-th smallest key and the
// tie counts are needed, never the order.
//
// Keys are the 31-bit patterns of non-negative fp32 values (|w| or the SNIP score); NaN is
// canonicalised to the largest key (torch.sort / topk place NaN last).  Three digits:
//   pass 0: bits 30..19 (exponent + 4 mantissa bits, 4096 bins) over ALL alive keys
//   pass 1: bits 18..7  (4096 bins)   pass 2: bits 6..0 (128 bins)
// After pass 0 the bucket that holds rank k is known.  If it fits the candidate buffer the
// second full-data pass gathers its (key, position) pairs ("collect" mode) and passes 1/2
// run on that small list; otherwise passes 1 and 2 re-stream the full data ("histogram"
// mode, degenerate inputs such as a constant tensor).  Per-CTA shared-memory histograms
// are flushed with one global atomic per non-empty bin; the last CTA to finish (ticket
// counter) scans the histogram and advances the select state, so no host round trip and
// no extra launch is needed between passes.  The staged entry points (hist / scan
// separately) exist so that a parameter-sharded multi-GPU select can all-reduce the
// histogram between them.
#include "common.cuh"
#include "comm.cuh"
#include "emit_body.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace b200p {

constexpr int kScanThreads = kThreads;           // the scan runs inside the last CTA
constexpr int kBinsPerThread = kHistBins / kScanThreads;   // 16

// ---- per-thread view of one chunk ------------------------------------------------------
// VEC path:    slot i = 4*j+q  <->  element 4*(j*kThreads+tid)+q
// scalar path: slot i          <->  element i*kThreads+tid
template <bool VEC>
__device__ __forceinline__ int slot_element(int i) {
    return VEC ? 4 * ((i >> 2) * kThreads + threadIdx.x) + (i & 3) : i * kThreads + threadIdx.x;
}

// Raw registers of one chunk (vector path): 4 float4 + the 16-bit alive bitmap.  Loaded one
// iteration ahead of its use so that the histogram / compare phase of the current chunk overlaps
// the loads of the next one.
struct ChunkRegs {
    float4 v[kVecPerThread];
    uint32_t alive;
    bool vec;          // false: partial or unaligned chunk, handled by the scalar path on demand
};

__device__ __forceinline__ void prefetch_chunk(const float* __restrict__ src, const uint32_t* __restrict__ mask_chunk,
                                               int n, int vec_ok, ChunkRegs& r) {
    r.vec = vec_ok && n == kChunk;
    r.alive = 0xFFFFu;
    if (r.vec) {
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) r.v[j] = ld_nc_f4(src + 4 * (j * kThreads + threadIdx.x));
        if (mask_chunk) {
            r.alive = 0;
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j)
                r.alive |= nibble_of(__ldg(mask_chunk + vec_word_index(j))) << (4 * j);
        }
    }
}

// 16 keys + alive bitmap of this thread for the chunk held in `r` (or loaded here, scalar path)
__device__ __forceinline__ void chunk_keys(const ChunkRegs& r, const float* __restrict__ src,
                                           const uint32_t* __restrict__ mask_chunk, int n,
                                           uint32_t (&key)[16], uint32_t& alive) {
    if (r.vec) {
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) {
            key[4 * j + 0] = key_of(r.v[j].x); key[4 * j + 1] = key_of(r.v[j].y);
            key[4 * j + 2] = key_of(r.v[j].z); key[4 * j + 3] = key_of(r.v[j].w);
        }
        alive = r.alive;
    } else {
        alive = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int e = i * kThreads + threadIdx.x;
            key[i] = 0;
            if (e < n) {
                key[i] = key_of(src[e]);
                uint32_t bit = 1u;
                if (mask_chunk) bit = (__ldg(mask_chunk + (e >> 5)) >> (e & 31)) & 1u;
                alive |= bit << i;
            }
        }
    }
}

// ---- scan of the global histogram by one CTA -------------------------------------------
// Finds the bin that holds rank st->k, advances the state, clears the histogram.
__device__ void scan_and_advance(int pass, unsigned long long* __restrict__ hist, SelState* __restrict__ st,
                                 long long cand_capacity) {
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    __shared__ unsigned long long s_total;
    const int tid = threadIdx.x;
    const int bins = digit_bins(pass);
    unsigned long long local[kBinsPerThread];
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) {
        const int b = tid * kBinsPerThread + i;
        // volatile: the counts were produced by other CTAs' atomics (L2), never cached in L1
        local[i] = b < bins ? ((volatile unsigned long long*)hist)[b] : 0ull;
        sum += local[i];
    }
    // block exclusive scan of `sum`
    unsigned long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        unsigned long long v = tid < kScanThreads / 32 ? s_warp[tid] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (tid >= o) inc += u;
        }
        if (tid < kScanThreads / 32) s_warp[tid] = inc - v;   // exclusive warp offsets
        if (tid == kScanThreads / 32 - 1) s_total = inc;
    }
    __syncthreads();
    unsigned long long running = s_warp[tid >> 5] + (incl - sum);
    const unsigned long long total = s_total;
    const unsigned long long k = st->k;
    const uint32_t prefix = st->prefix;
    __syncthreads();   // everyone has read the state before anyone writes it
    if (pass == 0 && tid == 0) st->n_valid = total;
    const int shift = pass == 0 ? 19 : pass == 1 ? 7 : 0;
    bool found = false;
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) {
        const int b = tid * kBinsPerThread + i;
        if (!found && b < bins && local[i] != 0 && running < k && k <= running + local[i]) {
            found = true;
            st->n_less += running;
            st->k = k - running;
            st->bucket_count = local[i];
            const uint32_t np = prefix | ((uint32_t)b << shift);
            st->prefix = np;
            if (pass == 0) {
                st->collect = (st->allow_collect && local[i] <= (unsigned long long)cand_capacity) ? 1u : 0u;
                st->cand_count = 0;
            }
            if (pass == 2) {
                st->thr_key = np;
                st->threshold = key_to_float(np);
                st->n_equal = local[i];
                st->quota = k - running;
                st->need_ties = (st->mode == B200P_MODE_EXACT_K && (k - running) < local[i]) ? 1u : 0u;
                st->tie_chunk = -1;
                st->tie_resid = 0;
                st->tie_seen = 0;
            }
        }
        running += local[i];
    }
    if (tid == 0 && (k == 0 || k > total)) {
        // rank outside the key set (host error): report the largest key so that nothing
        // above it survives a strict compare; n_equal = 0 flags the condition.
        st->prefix = kNanKey; st->thr_key = kNanKey; st->threshold = key_to_float(kNanKey);
        st->bucket_count = 0; st->n_equal = 0; st->quota = 0; st->need_ties = 0; st->collect = 0;
        st->tie_chunk = -1;
    }
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) {
        const int b = tid * kBinsPerThread + i;
        if (b < kHistBins) hist[b] = 0ull;
    }
}

// ---- candidate append with per-warp staging -------------------------------------------------
// A global atomic on the one shared counter per warp per chunk serialises in L2 (tens of thousands
// of same-address atomics per sweep).  Each warp stages its matches in shared memory and reserves
// global space once per ~128 candidates instead (sweep_loop below).
constexpr int kStage = 256;                       // staged candidates per warp

// ---- select state initialisation ----------------------------------------------------------
__device__ __forceinline__ void init_state(SelState* st, unsigned long long k, uint32_t mode, uint32_t allow_collect,
                                           uint32_t miss = 0u) {
    SelState s;
    memset(&s, 0, sizeof(s));
    s.k = k; s.k_request = k; s.mode = mode; s.allow_collect = allow_collect; s.tie_chunk = -1;
    s.miss = miss;           // never transiently cleared: other CTAs of the finish kernel may be reading it
    *st = s;
}

struct PassArgs {
    const int32_t* chunk_n;
    ChunkTab key_tab;
    const uint32_t* old_mask;
    unsigned long long* hist;
    SelState* st;
    uint32_t* cand_key;
    uint32_t* cand_pos;
    unsigned int* ticket;
    long long cand_capacity;
    uint32_t* prov;          // nullable: provisional packed mask (alive && key >= base) written by the sweep
    int64_t c_begin, c_end;
    int vec_ok, fuse_scan;
    // fused initialisation (pass 0 of b200p_select_kth): the last CTA resets the state before its scan
    int fuse_init;
    unsigned long long k;
    uint32_t mode, allow_collect;
    uint32_t comm_seq;       // != 0: sharded select, the last CTA of the bracket sweep all-reduces {fine histogram, below}
};

// ---- slim sweep used by exact pass 1 and by the bracket pass of the sampled select ------------
__device__ __forceinline__ float sel16(const float4 (&v)[kVecPerThread], int i) {
    const int j = i >> 2, q = i & 3;
    const float4 t = j == 0 ? v[0] : j == 1 ? v[1] : j == 2 ? v[2] : v[3];
    return q == 0 ? t.x : q == 1 ? t.y : q == 2 ? t.z : t.w;
}

// Sweeps chunks [c_begin, c_end): alive keys in [base, base + span) are "inside": they are counted in
// the fine histogram s_hist[(key - base) >> fine_shift] and (collect) appended to the candidate
// buffer through a per-warp staging area.  Returns this thread's count of alive keys below `base`.
// canon: the range reaches the inf/NaN buckets, keys must be canonicalised: element-wise path.
// Two chunks are loaded per iteration (128 B in flight per thread, 4 CTAs per SM) so that the sweep
// is bound by HBM and not by the latency of one 64-byte load per thread.
// per-warp candidate staging (one instance per kernel, shared by both instantiations of sweep_loop)
struct SweepStage {
    uint32_t key[kThreads / 32][kStage];
    uint32_t pos[kThreads / 32][kStage];
    int wcount[kThreads / 32];
};
__device__ __forceinline__ SweepStage& sweep_stage() { __shared__ SweepStage s_stage; return s_stage; }

template <bool NANP>
__device__ __forceinline__ unsigned long long sweep_loop(const PassArgs& a, uint32_t base, uint32_t span, int fine_shift,
                                                         bool collect, bool canon, uint32_t* __restrict__ s_hist) {
    SweepStage& stg = sweep_stage();
    uint32_t (&sg_key)[kThreads / 32][kStage] = stg.key;
    uint32_t (&sg_pos)[kThreads / 32][kStage] = stg.pos;
    int (&s_wcount)[kThreads / 32] = stg.wcount;
    SelState* __restrict__ st = a.st;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_wcount[warp] = 0;
    __syncwarp();
    unsigned long long below = 0;
    unsigned int nan_cnt = 0;                 // alive NaN keys: they sort last but SNIP_STRICT prunes them, the kept count needs them
    if (blockIdx.x == 0 && threadIdx.x == 0) st->prov_ok = (a.prov != nullptr && collect) ? 1u : 0u;

    // one matching key: fine histogram + staged append (order inside the buffer is irrelevant)
    auto take = [&](uint32_t key, uint32_t pos) {
        atomicAdd(&s_hist[(key - base) >> fine_shift], 1u);
        if (!collect) return;
        const int off = atomicAdd(&s_wcount[warp], 1);
        if (off < kStage) { sg_key[warp][off] = key; sg_pos[warp][off] = pos; }
        else {
            const uint32_t g = atomicAdd(&st->cand_count, 1u);
            if ((long long)g < a.cand_capacity) { a.cand_key[g] = key; a.cand_pos[g] = pos; }
        }
    };
    auto flush = [&](int n) {
        if (n <= 0) return;
        uint32_t gbase = 0;
        if (lane == 0) gbase = atomicAdd(&st->cand_count, (uint32_t)n);
        gbase = __shfl_sync(0xFFFFFFFFu, gbase, 0);
        for (int i = lane; i < n; i += 32)
            if ((long long)(gbase + i) < a.cand_capacity) { a.cand_key[gbase + i] = sg_key[warp][i]; a.cand_pos[gbase + i] = sg_pos[warp][i]; }
        __syncwarp();
        if (lane == 0) s_wcount[warp] = 0;
        __syncwarp();
    };
    // vector chunk held in registers.  Per key: AND, two subtracts, two funnel shifts that push the sign
    // bits of (k - base) and (k - base - span) into two 16-bit masks (first key ends up in bit 15).
    // k >= base  =>  k - base < 2^31, so the sign of (k - base - span) is the unsigned compare; keys
    // below `base` are removed from the inside-mask afterwards.
    // SNIP_STRICT prunes NaN scores (NaN > thr is false, train.py:316) although they sort last: their provisional bits
    // must be clear, the patching emit never sees them.  One max per key finds the (rare) threads that hold one.
    // NANP (SNIP_STRICT): NaN keys sort last but are pruned; the magnitude select (EXACT_K) keeps them like any large key and
    // skips the per-key max.  The keys are walked LAST TO FIRST, so that key i ends up in bit i of the 16-bit masks: memory
    // order, the alive bitmap and the mask nibbles need no bit reversal.
    constexpr bool nan_pruned = NANP;
    auto do_vec = [&](const float4 (&v)[kVecPerThread], uint32_t alive, uint32_t pos0) {
        uint32_t lt = 0, in = 0, mx = 0;
#pragma unroll
        for (int j = kVecPerThread - 1; j >= 0; --j) {
            const float f[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
            for (int q = 3; q >= 0; --q) {
                const uint32_t key = __float_as_uint(f[q]) & 0x7FFFFFFFu;
                const uint32_t d = key - base;
                if (NANP) mx = max(mx, key);
                lt = __funnelshift_l(d, lt, 1);
                in = __funnelshift_l(d - span, in, 1);
            }
        }
        lt &= alive;
        below += __popc(lt);
        if (a.prov) {
            // provisional mask: alive keys at or above the bracket base stay set; the emit only patches the
            // candidates afterwards instead of re-reading every key
            uint32_t keep = alive & ~lt & 0xFFFFu;
            if (NANP && mx > 0x7F800000u) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if ((__float_as_uint(sel16(v, i)) & 0x7FFFFFFFu) > 0x7F800000u) { nan_cnt += (keep >> i) & 1u; keep &= ~(1u << i); }
            }
            uint32_t* pw = a.prov + (size_t)(pos0 >> 12) * kWordsPerChunk;
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                const uint32_t word = gather_nibbles((keep >> (4 * j)) & 0xFu);        // key 4j+q sits at bit 4j+q
                if ((threadIdx.x & 7) == 0) pw[vec_word_index(j)] = word;
            }
        }
        uint32_t match = in & ~lt & alive & 0xFFFFu;
        while (match) {                               // divergent, rare (~1-2 % of the keys)
            const int i = __ffs(match) - 1;           // index of the key inside the thread's 16
            match &= match - 1;
            take(__float_as_uint(sel16(v, i)) & 0x7FFFFFFFu, pos0 + (uint32_t)slot_element<true>(i));
        }
    };
    // partial, unaligned or canonicalised chunk: element-wise
    auto do_scalar = [&](const float* __restrict__ src, const uint32_t* __restrict__ mchunk, int n, uint32_t pos0) {
        uint32_t* pw = a.prov ? a.prov + (size_t)(pos0 >> 12) * kWordsPerChunk : nullptr;
        for (int it = 0; it < kChunk / kThreads; ++it) {                  // warp-uniform trip count: one mask word per warp and step
            const int e = it * kThreads + threadIdx.x;
            bool alive = e < n;
            if (alive && mchunk) alive = (__ldg(mchunk + (e >> 5)) >> (e & 31)) & 1u;
            const uint32_t k = alive ? key_of(src[e]) : 0u;
            if (pw) {
                const uint32_t word = __ballot_sync(0xFFFFFFFFu, alive && k >= base && !(nan_pruned && k == kNanKey));
                if (lane == 0) pw[e >> 5] = word;
                if (alive && nan_pruned && k == kNanKey) ++nan_cnt;
            }
            if (alive) {
                if (k < base) ++below;
                else if (k - base < span) take(k, pos0 + (uint32_t)e);
            }
        }
    };
    // alive bitmap in do_vec's bit order (key i -> bit i)
    auto load_alive = [&](const uint32_t* __restrict__ mchunk) {
        uint32_t alive = 0xFFFFu;
        if (mchunk) {
            alive = 0;
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) alive |= nibble_of(__ldg(mchunk + vec_word_index(j))) << (4 * j);
        }
        return alive;
    };

    const bool vec_ok = a.vec_ok && !canon;
    // The chunk's address comes from a table: table load -> data load is a two-level dependent chain, and ncu showed both
    // levels exposed in every iteration (a quarter of all stall samples on the first use of the pointer and of the data).
    // The table entries of the NEXT iteration are requested before this iteration's data loads.
    const int64_t step = 2 * (int64_t)gridDim.x;
    const float* nsrc0 = nullptr; const float* nsrc1 = nullptr; int nn0 = 0, nn1 = 0;
    {
        const int64_t f0 = a.c_begin + blockIdx.x, f1 = f0 + gridDim.x;
        if (f0 < a.c_end) { nsrc0 = chunk_ptr<const float>(a.key_tab, f0); nn0 = __ldg(a.chunk_n + f0); }
        if (f1 < a.c_end) { nsrc1 = chunk_ptr<const float>(a.key_tab, f1); nn1 = __ldg(a.chunk_n + f1); }
    }
    for (int64_t c0 = a.c_begin + blockIdx.x; c0 < a.c_end; c0 += step) {
        const int64_t c1 = c0 + gridDim.x;
        const bool has1 = c1 < a.c_end;
        const float* src0 = nsrc0; const int n0 = nn0;
        const float* src1 = has1 ? nsrc1 : nullptr; const int n1 = has1 ? nn1 : 0;
        if (c0 + step < a.c_end) { nsrc0 = chunk_ptr<const float>(a.key_tab, c0 + step); nn0 = __ldg(a.chunk_n + c0 + step); }
        if (c1 + step < a.c_end) { nsrc1 = chunk_ptr<const float>(a.key_tab, c1 + step); nn1 = __ldg(a.chunk_n + c1 + step); }
        const uint32_t* m0 = a.old_mask ? a.old_mask + c0 * kWordsPerChunk : nullptr;
        const uint32_t* m1 = (a.old_mask && has1) ? a.old_mask + c1 * kWordsPerChunk : nullptr;
        const bool vec0 = vec_ok && n0 == kChunk, vec1 = has1 && vec_ok && n1 == kChunk;
        float4 v0[kVecPerThread], v1[kVecPerThread];
        uint32_t al0 = 0xFFFFu, al1 = 0xFFFFu;
        if (vec0) {
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) v0[j] = ld_nc_f4(src0 + 4 * (j * kThreads + threadIdx.x));
            al0 = load_alive(m0);
        }
        if (vec1) {
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) v1[j] = ld_nc_f4(src1 + 4 * (j * kThreads + threadIdx.x));
            al1 = load_alive(m1);
        }
        if (vec0) do_vec(v0, al0, (uint32_t)(c0 * kChunk)); else do_scalar(src0, m0, n0, (uint32_t)(c0 * kChunk));
        if (has1) { if (vec1) do_vec(v1, al1, (uint32_t)(c1 * kChunk)); else do_scalar(src1, m1, n1, (uint32_t)(c1 * kChunk)); }
        __syncwarp();
        if (collect) {
            // Broadcast lane 0's view of the fill level: under independent thread scheduling a lane may
            // run ahead into the next chunk and bump the counter before a slower lane has read it, and a
            // non-uniform decision here would leave the warp split across different barriers.
            const int filled = __shfl_sync(0xFFFFFFFFu, s_wcount[warp], 0);
            if (filled >= kStage / 2) flush(filled < kStage ? filled : kStage);
        }
    }
    __syncwarp();
    if (collect) {
        const int filled = __shfl_sync(0xFFFFFFFFu, s_wcount[warp], 0);
        flush(filled < kStage ? filled : kStage);
    }
    if (nan_cnt) atomicAdd(a.hist + kHistBins + 2, (unsigned long long)nan_cnt);      // rare
    return below;
}

// ---- full-data pass --------------------------------------------------------------------
// PASS 0: histogram digit 0 of every alive key.
// PASS 1: keys whose digit 0 matches the chosen bucket are histogrammed by digit 1 and, in collect
//         mode, appended to the candidate buffer as (key, position) in the same sweep.
// PASS 2: histogram digit 2 of the keys matching the 24 fixed bits — over the candidate buffer in
//         collect mode, over the full data otherwise.
template <int PASS>
__device__ __forceinline__ void pass_body(const PassArgs& a, uint32_t* __restrict__ s_hist) {
    SelState* __restrict__ st = a.st;
    uint32_t prefix = 0, collect = 0;
    if (PASS > 0) { prefix = st->prefix; collect = st->collect; }
    const int bins = digit_bins(PASS);
    for (int b = threadIdx.x; b < bins; b += kThreads) s_hist[b] = 0;
    __syncthreads();
    const uint32_t pmask = prefix_mask_before(PASS);
    const int lane = threadIdx.x & 31;

    if (PASS == 1) {
        // keys of the chosen 12-bit bucket: histogram of their next 12 bits, plus the candidate append
        if (st->mode == B200P_MODE_SNIP_STRICT) sweep_loop<true>(a, prefix, 1u << 19, 7, collect != 0, (prefix >> 19) >= 0xFF0u, s_hist);
        else                                    sweep_loop<false>(a, prefix, 1u << 19, 7, collect != 0, (prefix >> 19) >= 0xFF0u, s_hist);
    } else if (PASS == 2 && collect) {
        const uint32_t n = st->cand_count;
        for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
            const uint32_t key = a.cand_key[i];
            if ((key & pmask) == prefix) atomicAdd(&s_hist[key & 0x7Fu], 1u);
        }
    } else {
        int64_t c = a.c_begin + blockIdx.x;
        if (c < a.c_end) {
            const float* src = chunk_ptr<const float>(a.key_tab, c);
            int n = __ldg(a.chunk_n + c);
            ChunkRegs cur;
            prefetch_chunk(src, a.old_mask ? a.old_mask + c * kWordsPerChunk : nullptr, n, a.vec_ok, cur);
            while (true) {
                const int64_t cn = c + gridDim.x;
                const bool more = cn < a.c_end;
                const float* srcn = nullptr; int nn = 0;
                ChunkRegs nxt; nxt.vec = false; nxt.alive = 0;
                if (more) {
                    srcn = chunk_ptr<const float>(a.key_tab, cn);
                    nn = __ldg(a.chunk_n + cn);
                    prefetch_chunk(srcn, a.old_mask ? a.old_mask + cn * kWordsPerChunk : nullptr, nn, a.vec_ok, nxt);
                }
                uint32_t key[16], alive;
                chunk_keys(cur, src, a.old_mask ? a.old_mask + c * kWordsPerChunk : nullptr, n, key, alive);
                if (PASS == 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if ((alive >> i) & 1u) atomicAdd(&s_hist[key[i] >> 19], 1u);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (((alive >> i) & 1u) && ((key[i] & pmask) == prefix)) atomicAdd(&s_hist[digit_of(key[i], PASS)], 1u);
                }
                if (!more) break;
                c = cn; src = srcn; n = nn; cur = nxt;
            }
        }
    }
    __syncthreads();
    // flush the CTA histogram: one global atomic per non-empty bin
    for (int b = threadIdx.x; b < bins; b += kThreads) {
        const uint32_t v = s_hist[b];
        if (v) atomicAdd(a.hist + b, (unsigned long long)v);
    }
}

// measurement aid (b200p_select_last_trace): globaltimer stamps of the last sample / sweep kernel's last CTA
//   [0..3] sample: tail entered, bins staged, bracket chosen, done   [4..7] sweep: the same   [8] sweep: first CTA start
//   [9] sweep: last CTA's flush issued
__device__ unsigned long long g_sel_stamps[16];
__device__ __forceinline__ unsigned long long sel_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void sel_stamp(int i) { if (threadIdx.x == 0) g_sel_stamps[i] = sel_globaltimer(); }

// ---- histogram flush and last-CTA read-back ----------------------------------------------------
// Measured (tools/probes/hist_flush_probe.cu, hist_tail_probe.cu): 592 CTAs that each add their ~600-800 non-zero bins
// of a shared histogram into ONE global histogram at the end of a balanced sweep need 8-16 us for those atomics to
// drain (same-line serialisation in L2 plus ~100 atomics/ns chip-wide), and the last CTA then spent 2.5 us reading 16
// bins per thread at a 128-byte stride (one sector per lane) and 1.5 us clearing them the same way.  ncu showed it as
// sm__cycles_active.max - .min = 25 k cycles on every sample / sweep kernel.
//  * tried: clusters of 8 / 2 CTAs that sum their shared histograms through DSMEM before the global atomics (8x / 2x
//    fewer).  Clusters of 8 strand SMs on the 16/18/20-SM GPCs: the streaming body of the sweep ran 20 % slower; clusters
//    of 2 made no measurable difference.  Removed again;
//  * tried: an L2 access-policy window (persisting) over the histogram copies for these launches: no change (tail staged
//    in 4.2 against 4.0 us), so the cold part of the tail is not the histogram data;
//  * the global histogram has kHistReplicas copies, CTA i adds into copy i % kHistReplicas (4x fewer atomics per line);
//  * the last CTA reads the bins back coalesced into shared memory, summing and re-zeroing the copies on the way, and
//    clears copy 0 coalesced.
// track: the kernel does not know beforehand which bins it touches (the sample histogram): the range of non-zero bins is
// kept in two scalar counters behind the bins (both grow from zero: kHistBins - lowest bin, highest bin + 1)
constexpr int kExtraRangeLo = 3, kExtraRangeHi = 4;
__device__ __forceinline__ void flush_hist(const uint32_t* s_hist, unsigned long long* __restrict__ hist, bool track) {
    unsigned long long* dst = hist + (size_t)(blockIdx.x & (kHistReplicas - 1)) * kHistStride;
    int lo = kHistBins, hi = 0;
    for (int b = threadIdx.x; b < kHistBins; b += kThreads) {
        const uint32_t v = s_hist[b];
        if (v) { atomicAdd(dst + b, (unsigned long long)v); lo = min(lo, b); hi = b + 1; }
    }
    if (track) {
        lo = __reduce_min_sync(0xFFFFFFFFu, lo); hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        if ((threadIdx.x & 31) == 0 && hi > 0) {
            atomicMax(hist + kHistBins + kExtraRangeLo, (unsigned long long)(kHistBins - lo));
            atomicMax(hist + kHistBins + kExtraRangeHi, (unsigned long long)hi);
        }
    }
}

// The last CTA works alone, and what it reads is usually not in L2 any more (the sweep streamed hundreds of MB through it
// since the previous build zeroed these lines): reading all kHistReplicas x 32 KB took 7-16 us at the memory-level
// parallelism of one SM (b200p_select_last_trace).  Only the bins [lo, hi) that can be non-zero are read, summed, zeroed.
// Loads first, stores afterwards: a store into `hist` between the loads orders every later load behind it.
__device__ __forceinline__ void zero_hist_copies(unsigned long long* __restrict__ hist, int lo, int hi) {
#pragma unroll
    for (int r = 1; r < kHistReplicas; ++r)
#pragma unroll
        for (int i = 0; i < kBinsPerThread; ++i) {
            const int b = i * kThreads + threadIdx.x;
            if (b >= lo && b < hi) hist[(size_t)r * kHistStride + b] = 0ull;
        }
}
// sum of the copies -> copy 0 (the sharded select all-reduces copy 0 in place), copies 1.. zeroed again
__device__ __forceinline__ void collapse_hist(unsigned long long* __restrict__ hist, int lo, int hi) {
    unsigned long long sum[kBinsPerThread];
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) {
        const int b = i * kThreads + threadIdx.x;
        const bool in = b >= lo && b < hi;
        unsigned long long v = 0;
#pragma unroll
        for (int r = 0; r < kHistReplicas; ++r) v += in ? __ldcg(hist + (size_t)r * kHistStride + b) : 0ull;
        sum[i] = v;
    }
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) {
        const int b = i * kThreads + threadIdx.x;
        if (b >= lo && b < hi) hist[b] = sum[i];
    }
    zero_hist_copies(hist, lo, hi);
    __threadfence();
    __syncthreads();
}
// bins [lo, hi) of copy 0 (copies: of all copies, summed; copies 1.. zeroed again) -> s_stage, zero elsewhere
// (u32: a bin never holds 2^32 keys, positions are 32-bit)
__device__ __forceinline__ void stage_hist(unsigned long long* __restrict__ hist, uint32_t* s_stage, bool copies, int lo, int hi) {
    uint32_t sum[kBinsPerThread];
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) {
        const int b = i * kThreads + threadIdx.x;
        const bool in = b >= lo && b < hi;
        unsigned long long v = in ? __ldcg(hist + b) : 0ull;
#pragma unroll
        for (int r = 1; r < kHistReplicas; ++r) v += (in && copies) ? __ldcg(hist + (size_t)r * kHistStride + b) : 0ull;
        sum[i] = (uint32_t)v;
    }
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) s_stage[i * kThreads + threadIdx.x] = sum[i];
    if (copies) zero_hist_copies(hist, lo, hi);
    __syncthreads();
}

// true in exactly one CTA: the last one to arrive (all other CTAs' global writes are visible to it)
__device__ __forceinline__ bool last_cta_arrives(unsigned int* ticket) {
    __shared__ unsigned int s_ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const bool last = s_ticket == gridDim.x - 1;
    if (last) __threadfence();
    return last;
}

template <int PASS>
__global__ void __launch_bounds__(kThreads, 3)
k_select_pass(PassArgs a) {
    __shared__ uint32_t s_hist[kHistBins];
    pass_body<PASS>(a, s_hist);
    if (!a.fuse_scan) return;
    // the last CTA to finish scans the global histogram and advances the select state
    if (last_cta_arrives(a.ticket)) {
        if (PASS == 0 && a.fuse_init) {
            if (threadIdx.x == 0) init_state(a.st, a.k, a.mode, a.allow_collect);
            __syncthreads();
        }
        scan_and_advance(PASS, a.hist, a.st, a.cand_capacity);
        if (threadIdx.x == 0) *a.ticket = 0u;
    }
}

__global__ void k_select_scan(int pass, unsigned long long* hist, SelState* st, long long cand_capacity) {
    scan_and_advance(pass, hist, st, cand_capacity);
}

__global__ void k_select_init(SelState* st, unsigned long long* hist, unsigned int* ticket,
                              unsigned long long k, uint32_t mode, uint32_t allow_collect) {
    for (int b = threadIdx.x; b < kHistReplicas * kHistStride; b += blockDim.x) hist[b] = 0ull;
    if (threadIdx.x == 0) {
        init_state(st, k, mode, allow_collect);
        *ticket = 0u;
    }
}

// ---- tie resolution (EXACT_K with quota < n_equal) ---------------------------------------
// chunk_ties[c] = number of alive keys == threshold in chunk c.  Collect mode: every tie is in the
// candidate buffer (one atomic per tied candidate into a zeroed table); otherwise re-stream the keys.
__device__ void tie_count_vals(const int32_t* __restrict__ chunk_n, ChunkTab key_tab, const uint32_t* __restrict__ old_mask,
                               uint32_t need_ties, uint32_t thr, uint32_t collect, uint32_t cand_count, const uint32_t* __restrict__ cand_key,
                               const uint32_t* __restrict__ cand_pos, uint32_t* __restrict__ chunk_ties,
                               int64_t c_begin, int64_t c_end, int vec_ok) {
    if (!need_ties) return;
    if (collect) {
        const uint32_t n = cand_count;
        for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
            if (cand_key[i] == thr) atomicAdd(&chunk_ties[cand_pos[i] >> 12], 1u);
        return;
    }
    __shared__ int s_cnt;
    for (int64_t c = c_begin + blockIdx.x; c < c_end; c += gridDim.x) {
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        const float* src = chunk_ptr<const float>(key_tab, c);
        const int n = __ldg(chunk_n + c);
        const uint32_t* mchunk = old_mask ? old_mask + c * kWordsPerChunk : nullptr;
        ChunkRegs r;
        prefetch_chunk(src, mchunk, n, vec_ok, r);
        uint32_t key[16], alive;
        chunk_keys(r, src, mchunk, n, key, alive);
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) cnt += (((alive >> i) & 1u) && key[i] == thr) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        if (threadIdx.x == 0) chunk_ties[c] = (uint32_t)s_cnt;
        __syncthreads();
    }
}
__device__ __forceinline__ void tie_count_body(const int32_t* __restrict__ chunk_n, ChunkTab key_tab, const uint32_t* __restrict__ old_mask,
                                               const SelState* __restrict__ st, const uint32_t* __restrict__ cand_key,
                                               const uint32_t* __restrict__ cand_pos, uint32_t* __restrict__ chunk_ties,
                                               int64_t c_begin, int64_t c_end, int vec_ok) {
    tie_count_vals(chunk_n, key_tab, old_mask, st->need_ties, st->thr_key, st->collect, st->cand_count, cand_key, cand_pos, chunk_ties,
                   c_begin, c_end, vec_ok);
}
__global__ void __launch_bounds__(kThreads)
k_tie_count(const int32_t* __restrict__ chunk_n, ChunkTab key_tab, const uint32_t* __restrict__ old_mask,
            const SelState* __restrict__ st, const uint32_t* __restrict__ cand_key, const uint32_t* __restrict__ cand_pos,
            uint32_t* __restrict__ chunk_ties, int64_t c_begin, int64_t c_end, int vec_ok) {
    tie_count_body(chunk_n, key_tab, old_mask, st, cand_key, cand_pos, chunk_ties, c_begin, c_end, vec_ok);
}
// One CTA (any size up to 1024 threads): walk chunk_ties[c_begin, c_end) in order, find where the
// (quota - tie_offset)-th tie falls, clear the table again.
// d_counts (nullable): per-rank tie counts gathered from all ranks; the ties owned by the
// n_before lower ranks are added to tie_offset on the device (no host round trip).
__device__ void tie_scan_body(SelState* __restrict__ st, uint32_t* __restrict__ chunk_ties, int64_t c_begin, int64_t c_end,
                              unsigned long long tie_offset, const unsigned long long* __restrict__ d_counts, int n_before,
                              unsigned long long* s_part /* [blockDim.x] */) {
    if (!st->need_ties) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (d_counts) for (int r = 0; r < n_before; ++r) tie_offset += d_counts[r];
    const int64_t n = c_end - c_begin;
    const int64_t per = (n + nt - 1) / nt;
    const int64_t lo = c_begin + tid * per < c_end ? c_begin + tid * per : c_end;
    const int64_t hi = lo + per < c_end ? lo + per : c_end;
    unsigned long long sum = 0;
    for (int64_t c = lo; c < hi; ++c) sum += chunk_ties[c];
    s_part[tid] = sum;
    __syncthreads();
    __shared__ unsigned long long s_tie_total;
    if (tid == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < nt; ++i) { unsigned long long v = s_part[i]; s_part[i] = run; run += v; }
        s_tie_total = run;
    }
    __syncthreads();
    const unsigned long long quota = st->quota;
    // ties still to prune inside this chunk range
    const long long target = (long long)quota - (long long)tie_offset;
    unsigned long long before = s_part[tid];
    const unsigned long long total = s_tie_total;
    __syncthreads();
    if (target <= 0) {
        if (tid == 0) { st->tie_chunk = c_begin; st->tie_resid = 0; st->tie_seen = 0; }
    } else if ((unsigned long long)target > total) {
        if (tid == 0) { st->tie_chunk = c_end; st->tie_resid = 0; st->tie_seen = total; }   // every local tie pruned
    } else {
        bool hit = false;
        for (int64_t c = lo; c < hi; ++c) {
            const unsigned long long v = chunk_ties[c];
            if (!hit && v != 0 && before < (unsigned long long)target && (unsigned long long)target <= before + v) {
                hit = true;
                st->tie_chunk = c;
                st->tie_resid = (uint32_t)((unsigned long long)target - before);
                st->tie_seen = before;
            }
            before += v;
        }
    }
    __syncthreads();
    for (int64_t c = lo; c < hi; ++c) chunk_ties[c] = 0;
}
// Few ties, all of them in the candidate buffer (the normal case: a handful of weights share the threshold's bit pattern):
// their positions go to a short list instead of the per-chunk table, and one CTA ranks the list.  The table walk above
// is one CTA striding through n_chunks counters three times — 75 us for ViT-L/16's 55 k chunks, 30 us for ResNet-152
// (tools/select_probe.py), for 6 tied keys.  tie_list[0] = count (left zero), positions from tie_list[1].
__device__ __forceinline__ void tie_list_gather(uint32_t thr, uint32_t cand_count, const uint32_t* __restrict__ cand_key,
                                                const uint32_t* __restrict__ cand_pos, uint32_t* __restrict__ tie_list) {
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < cand_count; i += gridDim.x * kThreads)
        if (__ldg(cand_key + i) == thr) {
            const uint32_t slot = atomicAdd(tie_list, 1u);
            if (slot < (uint32_t)kTieListCap) tie_list[1 + slot] = __ldg(cand_pos + i);
        }
}
// One CTA, after a grid-wide barrier: same outcome as tie_scan_body (lowest flat index first; ties owned by the n_before
// lower ranks count against the quota first).
// prov != nullptr: the CTA also clears the bits of the ties it prunes (the patch pass left every tie alone), and the state
// says "no tie handling left" (tie_chunk = c_begin, tie_resid = 0): nobody walks the tie chunk element by element any more
// (20 us on the rank that owns it: 128 dependent loads by one warp, the other ranks waiting at the final flag).
__device__ void tie_list_pick(SelState* __restrict__ st, uint32_t* __restrict__ tie_list, int64_t c_begin, int64_t c_end,
                              unsigned long long tie_offset, const unsigned long long* __restrict__ d_counts, int n_before,
                              uint32_t* s_pos /* [kTieListCap] */, uint32_t* __restrict__ prov = nullptr) {
    __shared__ int s_sel;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (d_counts) for (int r = 0; r < n_before; ++r) tie_offset += d_counts[r];
    uint32_t m = __ldcg(tie_list);
    if (m > (uint32_t)kTieListCap) m = kTieListCap;
    for (uint32_t i = tid; i < m; i += nt) s_pos[i] = __ldcg(tie_list + 1 + i);
    if (tid == 0) s_sel = -1;
    __syncthreads();
    const long long target = (long long)st->quota - (long long)tie_offset;
    if (prov) {
        const long long tgt = target < 0 ? 0 : target > (long long)m ? (long long)m : target;
        for (uint32_t i = tid; i < m; i += nt) {
            const uint32_t pi = s_pos[i];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < m; ++j) rank += s_pos[j] < pi ? 1u : 0u;                    // positions are distinct
            if ((long long)rank < tgt) atomicAnd(prov + (size_t)(pi >> 12) * kWordsPerChunk + ((pi & 4095u) >> 5), ~(1u << (pi & 31u)));
        }
        if (tid == 0) { st->tie_chunk = c_begin; st->tie_resid = 0; st->tie_seen = (unsigned long long)tgt; }
    } else if (target <= 0) {
        if (tid == 0) { st->tie_chunk = c_begin; st->tie_resid = 0; st->tie_seen = 0; }
    } else if ((unsigned long long)target > (unsigned long long)m) {
        if (tid == 0) { st->tie_chunk = c_end; st->tie_resid = 0; st->tie_seen = m; }           // every local tie pruned
    } else {
        for (uint32_t i = tid; i < m; i += nt) {
            const uint32_t pi = s_pos[i];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < m; ++j) rank += s_pos[j] < pi ? 1u : 0u;                    // positions are distinct
            if ((long long)rank == target - 1) s_sel = (int)i;
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t chunk = s_pos[s_sel] >> 12;
            uint32_t seen = 0;
            for (uint32_t j = 0; j < m; ++j) seen += (s_pos[j] >> 12) < chunk ? 1u : 0u;
            st->tie_chunk = (long long)chunk; st->tie_seen = seen; st->tie_resid = (uint32_t)(target - (long long)seen);
        }
    }
    __syncthreads();
    if (tid == 0) tie_list[0] = 0u;
}

__global__ void __launch_bounds__(1024)
k_tie_scan(SelState* __restrict__ st, uint32_t* __restrict__ chunk_ties, int64_t c_begin, int64_t c_end,
           unsigned long long tie_offset, const unsigned long long* __restrict__ d_counts, int n_before) {
    __shared__ unsigned long long s_part[1024];
    tie_scan_body(st, chunk_ties, c_begin, c_end, tie_offset, d_counts, n_before, s_part);
}

// One CTA: *out = sum of chunk_ties[c_begin, c_end) (0 when no tie resolution is needed).
__global__ void __launch_bounds__(1024)
k_tie_total(const SelState* __restrict__ st, const uint32_t* __restrict__ chunk_ties, int64_t c_begin, int64_t c_end,
            unsigned long long* __restrict__ out) {
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    unsigned long long sum = 0;
    if (st->need_ties)
        for (int64_t c = c_begin + threadIdx.x; c < c_end; c += blockDim.x) sum += chunk_ties[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(&s_sum, sum);
    __syncthreads();
    if (threadIdx.x == 0) *out = s_sum;
}

// =============================================================================================
// Sampled (bracketed) select — the default behind b200p_select_kth.
//
// The exact radix pass 0 pays one shared-memory atomic per key (2 cycles per lane: ~44 us for the
// 25.5 M keys of ResNet-50, 2.5x the HBM time).  Only the bucket that holds rank k matters, so:
//   S  k_select_sample   histogram of a 1/64 sample (and popcount of the old mask): the last CTA turns
//                        sample rank r = k*S/n_alive +- 4*sqrt(S) into a bracket of 1..8 twelve-bit buckets
//   A  k_select_bracket  ONE full-data sweep at memory speed: count the keys below the bracket in
//                        registers, histogram the (few %) keys inside it by their next 12 bits and append
//                        them to the candidate buffer; the last CTA verifies that rank k falls inside
//                        (else: miss) and narrows to a 1024-key window
//   B  k_select_finish   cooperative launch: histogram of the window over the candidates -> exact key,
//                        tie count and tie scan (EXACT_K).  On a miss (or an unusable sample) the same
//                        launch runs the exact 3-pass radix select with grid-wide barriers instead, so
//                        the result never depends on the sample.
// Three launches, no host round trip, one read of the keys.
// =============================================================================================
constexpr int kSampleSlotsPerChunk = 16;  // float4 sampled per 4096-key chunk (1/64 of the keys)
constexpr int kMaxBracketBuckets = 8;     // 8 * 2^19 keys >> 10 = 4096 fine bins
constexpr int kFineShift = 10;
constexpr int kWindow = 1 << kFineShift;

struct SampleArgs {
    const int32_t* chunk_n;
    ChunkTab key_tab;
    const uint32_t* old_mask;
    unsigned long long* hist;
    SelState* st;
    unsigned int* ticket;
    int64_t n_chunks;        // chunks sampled: [c_begin, c_begin + n_chunks)
    int64_t c_begin;         // first chunk of this rank's range (0 on one GPU)
    int vec_ok;
    unsigned long long k, n_total;
    uint32_t mode;
    uint32_t sigmas;         // bracket half-width in standard deviations of the sample rank (8; 12 for clustered granules)
    uint32_t comm_seq;       // != 0: parameter-sharded select, the last CTA all-reduces the histogram over the ranks first
    unsigned long long* cache;   // nullable: where the last CTA keeps a copy of the final sample histogram
    int slot_shift;          // k_select_sample: 2^slot_shift float4 sampled per chunk (sample_slot_shift)
};

// bracket for a new rank k from the cached sample histogram of the same keys: one CTA, no sampling, no collective
__global__ void __launch_bounds__(kThreads)
k_sample_from_cache(SampleArgs a, const unsigned long long* __restrict__ cache);

// exclusive prefix of this thread's 16 bins over the CTA (kScanThreads threads) and the grand total
__device__ __forceinline__ unsigned long long block_prefix16(const unsigned long long (&local)[kBinsPerThread],
                                                             unsigned long long* s_warp /*[9]*/, unsigned long long& total) {
    const int tid = threadIdx.x;
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) sum += local[i];
    unsigned long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    __syncthreads();
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        unsigned long long v = tid < kScanThreads / 32 ? s_warp[tid] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (tid >= o) inc += u;
        }
        if (tid < kScanThreads / 32) s_warp[tid] = inc - v;
        if (tid == kScanThreads / 32 - 1) s_warp[8] = inc;
    }
    __syncthreads();
    total = s_warp[8];
    return s_warp[tid >> 5] + (incl - sum);
}

__device__ __forceinline__ void clear_hist(unsigned long long* hist, int lo = 0, int hi = kHistBins) {
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) {                                                  // coalesced
        const int b = i * kThreads + threadIdx.x;
        if (b >= lo && b < hi) hist[b] = 0ull;
    }
    if (threadIdx.x < kHistExtra) hist[kHistBins + threadIdx.x] = 0ull;
}

// bracket of 1..8 buckets around the sample rank of k, from the sample histogram staged in s_stage
__device__ __forceinline__ void sample_pick(const SampleArgs& a, const uint32_t* s_stage, unsigned long long n_alive,
                                            unsigned long long* s_warp /*[9]*/, uint32_t* s_bkt /*[2]*/) {
    SelState* st = a.st;
    unsigned long long local[kBinsPerThread];
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) local[i] = s_stage[threadIdx.x * kBinsPerThread + i];
    unsigned long long S;
    unsigned long long running = block_prefix16(local, s_warp, S);
    if (threadIdx.x == 0) { s_bkt[0] = 0xFFFFFFFFu; s_bkt[1] = 0xFFFFFFFFu; init_state(st, a.k, a.mode, 1u); st->n_valid = n_alive; }
    __syncthreads();
    bool usable = S >= 1024 && a.k >= 1 && a.k <= n_alive;
    unsigned long long r_lo = 1, r_hi = 1;
    if (usable) {
        // sample rank of the population's k-th key: hypergeometric, sigma <= sqrt(S)/2; margin = 8 sigma_max + 2
        // (k <= N < 2^32 and S < 2^31 in every real case: the product fits 64 bits; the 128-bit division is ~1 us of one thread)
        const unsigned long long r = ((a.k >> 32) | (S >> 31)) ? (unsigned long long)(((__uint128_t)a.k * S + n_alive - 1) / n_alive)
                                                              : (a.k * S + n_alive - 1) / n_alive;
        // 8 sigma of the hypergeometric rank, sigma^2 <= S q (1-q), plus slack for tiny tails
        const double q = (double)a.k / (double)n_alive;
        const unsigned long long m = (unsigned long long)ceil((double)a.sigmas * sqrt((double)S * q * (1.0 - q))) + 16ull;
        r_lo = r > m ? r - m : 1ull;  if (r_lo < 1) r_lo = 1;
        r_hi = r + m < S ? r + m : S; if (r_hi < 1) r_hi = 1;
#pragma unroll
        for (int i = 0; i < kBinsPerThread; ++i) {
            const unsigned long long v = local[i];
            if (v != 0 && running < r_lo && r_lo <= running + v) s_bkt[0] = threadIdx.x * kBinsPerThread + i;
            if (v != 0 && running < r_hi && r_hi <= running + v) s_bkt[1] = threadIdx.x * kBinsPerThread + i;
            running += v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t lo = s_bkt[0], hi = s_bkt[1];
        const bool ok = usable && lo != 0xFFFFFFFFu && hi != 0xFFFFFFFFu && hi >= lo && hi - lo < (uint32_t)kMaxBracketBuckets &&
                        hi < 0xFF0u;      // the sweep does not canonicalise NaN keys: stay below the inf/NaN buckets
        st->sample_ok = ok ? 1u : 0u;
        st->miss = ok ? 0u : 1u;
        st->lo_bucket = lo; st->hi_bucket = hi;
        *a.ticket = 0u;
    }
}

// last CTA of a sample kernel.  comm: parameter-sharded select: the sample histograms (and alive counts) of all ranks are
// summed first, then every rank derives the same bracket.  Returns false if the collective failed.
__device__ __forceinline__ bool sample_tail(const SampleArgs& a, const CommDev* comm, unsigned long long* s_warp /*[9]*/, uint32_t* s_bkt /*[2]*/,
                                            uint32_t* s_stage /*[kHistBins]*/) {
    sel_stamp(0);
    // scalar loads first: their latency hides behind the staging loads
    const int lo = kHistBins - (int)((volatile unsigned long long*)a.hist)[kHistBins + kExtraRangeLo],
              hi = (int)((volatile unsigned long long*)a.hist)[kHistBins + kExtraRangeHi];
    bool comm_ok = true;
    if (comm && a.comm_seq) {
        collapse_hist(a.hist, lo, hi);
        comm_ok = comm_allreduce_hist(*comm, a.comm_seq, a.hist);
        stage_hist(a.hist, s_stage, false, 0, kHistBins);
    } else {
        stage_hist(a.hist, s_stage, true, lo, hi);
    }
    const unsigned long long n_alive = a.old_mask ? ((volatile unsigned long long*)a.hist)[kHistBins + 0] : a.n_total;
    sel_stamp(1);
    if (a.cache) {
        // keep the (global) sample histogram: another select over the SAME keys (a sparsity sweep) derives its bracket from
        // it without sampling again (B200P_OPT_REUSE_SAMPLE)
#pragma unroll
        for (int i = 0; i < kBinsPerThread; ++i) a.cache[i * kThreads + threadIdx.x] = s_stage[i * kThreads + threadIdx.x];
        if (threadIdx.x == 0) a.cache[kHistBins + 0] = n_alive;
    }
    sample_pick(a, s_stage, n_alive, s_warp, s_bkt);
    sel_stamp(2);
    if (comm && a.comm_seq) clear_hist(a.hist); else clear_hist(a.hist, lo, hi);
    sel_stamp(3);
    return comm_ok;
}

__global__ void __launch_bounds__(kThreads)
k_select_sample(SampleArgs a, CommDev comm) {
    __shared__ uint32_t s_hist[kHistBins];
    __shared__ unsigned long long s_warp[9];
    __shared__ unsigned long long s_alive;
    __shared__ uint32_t s_bkt[2];
    for (int b = threadIdx.x; b < kHistBins; b += kThreads) s_hist[b] = 0;
    if (threadIdx.x == 0) s_alive = 0;
    __syncthreads();
    const int64_t gtid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nthreads = (int64_t)gridDim.x * kThreads;
    if (a.old_mask) {
        unsigned long long cnt = 0;
        const int64_t words = a.n_chunks * kWordsPerChunk;
        const uint32_t* __restrict__ mw = a.old_mask + a.c_begin * kWordsPerChunk;
        for (int64_t w = gtid; w < words; w += nthreads) cnt += __popc(__ldg(mw + w));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_alive, cnt);
    }
    // sample slot s = (chunk, i), i < 16: elements 256*i + 4*phase .. +3 (one float4 out of 64), the
    // phase varying with chunk and i.  Four slots per thread are in flight at once: table loads,
    // then data loads, then the shared-memory atomics.
    const int64_t slots = a.n_chunks << a.slot_shift;
    const int slot_mask = (1 << a.slot_shift) - 1, slot_step = kSampleSlotsPerChunk >> a.slot_shift;
    for (int64_t s0 = gtid; s0 < slots; s0 += 4 * nthreads) {
        int n[4], e0[4]; const float* src[4]; int64_t cc[4]; bool on[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t sl = s0 + u * nthreads;
            on[u] = sl < slots;
            cc[u] = a.c_begin + (on[u] ? (sl >> a.slot_shift) : 0);
            const int i = (int)(sl & slot_mask) * slot_step + (int)(cc[u] & (slot_step - 1));      // thinned samples rotate through the 16 positions
            e0[u] = 256 * i + 4 * (int)((cc[u] + 5 * i) & 63);
            n[u] = __ldg(a.chunk_n + cc[u]);
            src[u] = chunk_ptr<const float>(a.key_tab, cc[u]);
        }
        float v[4][4]; uint32_t alive[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            on[u] = on[u] && e0[u] < n[u];
            alive[u] = 0xFu;
#pragma unroll
            for (int q = 0; q < 4; ++q) v[u][q] = 0.f;
            if (on[u]) {
                if (a.vec_ok && e0[u] + 3 < n[u]) {
                    const float4 t = ld_nc_f4(src[u] + e0[u]);
                    v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
                } else {
                    for (int q = 0; q < 4; ++q) if (e0[u] + q < n[u]) v[u][q] = src[u][e0[u] + q];
                }
                if (a.old_mask) alive[u] = (__ldg(a.old_mask + cc[u] * kWordsPerChunk + (e0[u] >> 5)) >> (e0[u] & 31)) & 0xFu;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!on[u]) continue;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (((alive[u] >> q) & 1u) && e0[u] + q < n[u]) atomicAdd(&s_hist[key_of(v[u][q]) >> 19], 1u);
        }
    }
    __syncthreads();
    flush_hist(s_hist, a.hist, true);
    if (threadIdx.x == 0 && s_alive) atomicAdd(a.hist + kHistBins + 0, s_alive);
    if (!last_cta_arrives(a.ticket)) return;
    const bool comm_ok = sample_tail(a, &comm, s_warp, s_bkt, s_hist);
    if (!comm_ok && threadIdx.x == 0) { a.st->sample_ok = 0u; a.st->miss = 1u; }
}

__global__ void __launch_bounds__(kThreads)
k_sample_from_cache(SampleArgs a, const unsigned long long* __restrict__ cache) {
    __shared__ uint32_t s_stage[kHistBins];
    __shared__ unsigned long long s_warp[9];
    __shared__ uint32_t s_bkt[2];
    for (int b = threadIdx.x; b < kHistBins; b += kThreads) s_stage[b] = (uint32_t)cache[b];
    const unsigned long long n_alive = a.old_mask ? cache[kHistBins + 0] : a.n_total;
    __syncthreads();
    sample_pick(a, s_stage, n_alive, s_warp, s_bkt);
}

// last CTA of a bracket sweep: verify that rank k is inside the bracket, narrow to a 1024-key window
__device__ __forceinline__ bool bracket_tail(const PassArgs& a, const CommDev* comm, uint32_t base, uint32_t span, unsigned long long* s_warp /*[9]*/,
                                             uint32_t* s_stage /*[kHistBins]*/) {
    SelState* __restrict__ st = a.st;
    sel_stamp(4);
    const int lo = 0, hi = (int)(span >> kFineShift);                 // the only fine bins a key inside the bracket can land in
    bool comm_ok = true;
    if (comm && a.comm_seq) {
        collapse_hist(a.hist, lo, hi);
        comm_ok = comm_allreduce_hist(*comm, a.comm_seq, a.hist);
    }
    // scalar loads first: their latency hides behind the staging loads
    const unsigned long long n_below = ((volatile unsigned long long*)a.hist)[kHistBins + 1];
    const unsigned long long n_nan = ((volatile unsigned long long*)a.hist)[kHistBins + 2];
    const unsigned long long k = st->k;
    stage_hist(a.hist, s_stage, !(comm && a.comm_seq), lo, hi);
    sel_stamp(5);
    unsigned long long local[kBinsPerThread];
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) local[i] = s_stage[threadIdx.x * kBinsPerThread + i];
    unsigned long long total_in;
    unsigned long long running = block_prefix16(local, s_warp, total_in);
    if (threadIdx.x == 0) st->pad_[1] = (uint32_t)n_nan;              // NaN keys pruned by the provisional mask
    const bool inside = k > n_below && k <= n_below + total_in && total_in <= (unsigned long long)a.cand_capacity;
    if (!inside) {
        if (threadIdx.x == 0) { st->miss = 1u; st->collect = 0u; st->cand_count = 0u; st->prov_ok = 0u; }
    } else {
        const unsigned long long kk = k - n_below;
#pragma unroll
        for (int i = 0; i < kBinsPerThread; ++i) {
            const unsigned long long v = local[i];
            if (v != 0 && running < kk && kk <= running + v) {
                st->n_less = n_below + running;
                st->k = kk - running;
                st->bucket_count = v;
                st->win_lo = base + ((uint32_t)(threadIdx.x * kBinsPerThread + i) << kFineShift);
                st->collect = 1u;
                st->passes_full = 1u;
            }
            running += v;
        }
    }
    if (threadIdx.x == 0) *a.ticket = 0u;
    sel_stamp(6);
    clear_hist(a.hist, lo, (comm && a.comm_seq) ? kHistBins : hi);
    sel_stamp(7);
    return comm_ok;
}

// ---- A: one sweep: count below the bracket, fine histogram + collect inside it -----------------
// The hot loop is issue-bound if written naively (16 keys per thread per chunk): it is kept to one
// AND, one subtract and two compares per key, builds 16-bit "below" / "inside" masks, applies the
// alive bitmap once per thread, and leaves everything that concerns the ~1 % matching keys (fine
// histogram, staging, position) to a divergent slow path that extracts the key by index.
__global__ void __launch_bounds__(kThreads, 4)
k_select_bracket(PassArgs a, CommDev comm) {
    __shared__ uint32_t s_hist[kHistBins];
    __shared__ unsigned long long s_warp[9];
    __shared__ unsigned long long s_below;
    SelState* __restrict__ st = a.st;
    if (!st->sample_ok) return;                      // the finish kernel runs the exact select instead
    const uint32_t lo_b = st->lo_bucket, hi_b = st->hi_bucket;
    const uint32_t base = lo_b << 19, span = (hi_b - lo_b + 1) << 19;
    for (int b = threadIdx.x; b < kHistBins; b += kThreads) s_hist[b] = 0;
    if (threadIdx.x == 0) s_below = 0;
    if (blockIdx.x == 0) sel_stamp(8);
    __syncthreads();
    // the state was initialised by the kernel before this one: the mode is uniform over the grid
    unsigned long long below = st->mode == B200P_MODE_SNIP_STRICT ? sweep_loop<true>(a, base, span, kFineShift, true, false, s_hist)
                                                                  : sweep_loop<false>(a, base, span, kFineShift, true, false, s_hist);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xFFFFFFFFu, below, o);
    if ((threadIdx.x & 31) == 0 && below) atomicAdd(&s_below, below);
    __syncthreads();
    flush_hist(s_hist, a.hist, false);
    if (threadIdx.x == 0 && s_below) atomicAdd(a.hist + kHistBins + 1, s_below);
    if (threadIdx.x == 0) atomicMax(&g_sel_stamps[9], sel_globaltimer());
    if (!last_cta_arrives(a.ticket)) return;
    const bool comm_ok = bracket_tail(a, &comm, base, span, s_warp, s_hist);
    if (!comm_ok && threadIdx.x == 0) { st->miss = 1u; st->collect = 0u; st->prov_ok = 0u; }
}

// One CTA: all-gather of the ranks' window histograms (a.hist[0..1024) holds this rank's), exact key of rank k, tie
// bookkeeping (every rank's count of that key -> rank_ties).  Leaves a.hist cleared.
__device__ __forceinline__ void sharded_gather_pick(const PassArgs& a, const CommDev& comm, uint32_t seq, unsigned long long* __restrict__ rank_ties,
                                                    unsigned long long* s_warp /*[9]*/, int* s_bin, uint32_t* s_stage /*[kWindow]*/) {
    SelState* __restrict__ st = a.st;
    const uint32_t win_lo = st->win_lo;
    const int tid = threadIdx.x, slot = seq & 1u;
    // push the own 1024 counts (4 per thread) into every window, wait for everybody's
    {
        uint4 v;
        v.x = (uint32_t)((volatile unsigned long long*)a.hist)[4 * tid + 0]; v.y = (uint32_t)((volatile unsigned long long*)a.hist)[4 * tid + 1];
        v.z = (uint32_t)((volatile unsigned long long*)a.hist)[4 * tid + 2]; v.w = (uint32_t)((volatile unsigned long long*)a.hist)[4 * tid + 3];
        for (int p = 0; p < comm.world; ++p)
            reinterpret_cast<uint4*>(comm.win[p] + comm.lay.gather)[(size_t)(slot * comm.world + comm.rank) * (kCommGatherWords / 4) + tid] = v;
    }
    const bool comm_ok = comm_signal_and_wait(comm, CH_GATHER, seq);
    const uint32_t* __restrict__ g = reinterpret_cast<const uint32_t*>(comm.win[comm.rank] + comm.lay.gather) + (size_t)slot * comm.world * kCommGatherWords;
    // coalesced: thread t sums bins t, t + 256, ... over the ranks into shared memory (16 consecutive bins per thread straight
    // from global memory are one 32-byte sector per lane and load: 8192 sector requests from 64 threads at 8 ranks)
    const unsigned long long k = st->k;
    {
        uint32_t sum[kWindow / kThreads];
#pragma unroll
        for (int i = 0; i < kWindow / kThreads; ++i) {
            uint32_t v = 0;
#pragma unroll
            for (int r = 0; r < kCommMaxWorld; ++r) v += r < comm.world ? __ldcg(g + (size_t)r * kCommGatherWords + i * kThreads + tid) : 0u;
            sum[i] = v;
        }
#pragma unroll
        for (int i = 0; i < kWindow / kThreads; ++i) s_stage[i * kThreads + tid] = sum[i];
        __syncthreads();
    }
    unsigned long long local[kBinsPerThread];
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) {
        const int b = tid * kBinsPerThread + i;
        local[i] = b < kWindow ? s_stage[b] : 0u;
    }
    unsigned long long total;
    unsigned long long running = block_prefix16(local, s_warp, total);
    if (tid == 0) *s_bin = -1;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kBinsPerThread; ++i) {
        const unsigned long long v = local[i];
        if (v != 0 && running < k && k <= running + v) {
            const int b = tid * kBinsPerThread + i;
            const uint32_t key = win_lo + (uint32_t)b;
            st->n_less += running;
            st->k = k - running;
            st->prefix = key; st->thr_key = key; st->threshold = key_to_float(key);
            st->n_equal = v; st->quota = k - running;
            st->need_ties = (st->mode == B200P_MODE_EXACT_K && (k - running) < v) ? 1u : 0u;
            st->tie_chunk = -1; st->tie_resid = 0; st->tie_seen = 0;
            *s_bin = b;
        }
        running += v;
    }
    __syncthreads();
    if (*s_bin < 0 || !comm_ok) {                    // cannot happen after a verified bracket; treated like a miss
        if (tid == 0) { st->miss = 1u; st->collect = 0u; st->prov_ok = 0u; }
    } else if (tid < kCommMaxWorld) {
        rank_ties[tid] = tid < comm.world ? (unsigned long long)__ldcg(g + (size_t)tid * kCommGatherWords + *s_bin) : 0ull;
    }
    __syncthreads();
    clear_hist(a.hist);
}

// ---- B (sharded): exact key from the window histograms of all ranks -----------------------------------------------
// Each rank histograms the candidates it collected (its chunk range only) over the 1024-key window the bracket tail
// chose; the last CTA all-GATHERS the 1024 counts of every rank (4 KB per rank), sums them, finds the key of rank k and,
// because it holds every rank's count of that key, also knows how many tied keys live in lower-ranked slices: the tie
// bookkeeping of SURVEY 8e needs no collective of its own.
__global__ void __launch_bounds__(kThreads)
k_sharded_finish(PassArgs a, CommDev comm, uint32_t seq, unsigned long long* __restrict__ rank_ties) {
    __shared__ uint32_t s_hist[kWindow];
    __shared__ unsigned long long s_warp[9];
    __shared__ int s_bin;
    SelState* __restrict__ st = a.st;
    if (st->miss) return;                            // same on every rank: the builder reruns the exact staged select
    for (int b = threadIdx.x; b < kWindow; b += kThreads) s_hist[b] = 0;
    __syncthreads();
    const uint32_t n = st->cand_count, win_lo = st->win_lo;
    const uint32_t stride = gridDim.x * kThreads;
    for (uint32_t i0 = blockIdx.x * kThreads + threadIdx.x; i0 < n; i0 += 8 * stride) {
        uint32_t kk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) kk[u] = i0 + u * stride < n ? __ldg(a.cand_key + i0 + u * stride) : 0xFFFFFFFFu;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t d = kk[u] - win_lo;
            if (i0 + u * stride < n && d < (uint32_t)kWindow) atomicAdd(&s_hist[d], 1u);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kWindow; b += kThreads) {
        const uint32_t v = s_hist[b];
        if (v) atomicAdd(a.hist + b, (unsigned long long)v);
    }
    if (!last_cta_arrives(a.ticket)) return;
    sharded_gather_pick(a, comm, seq, rank_ties, s_warp, &s_bin, s_hist);
    if (threadIdx.x == 0) *a.ticket = 0u;
}

// =============================================================================================
// Fused SNIP mask build (b200p_snip_mask_build): the score pass IS the sweep.
//   S' k_snip_sample<ACC,B>       scores of the 1/64 sample positions computed from W and the B gradient sets
//                                 (same slots, same bracket logic as k_select_sample)
//   A' k_snip_score_sweep<ACC,B>  SCORE = sum_b |W*G_b| written once (bit-identical to k_score_multi) and, while the
//                                 16 scores of a thread are still in registers, classified against the bracket:
//                                 below-count, fine histogram + candidates inside it, provisional mask word.  The
//                                 separate read of the scores by k_select_bracket (4 B/param and ~40 us of issue-bound
//                                 sweep for ResNet-50) disappears.
// then k_select_finish and the patching emit exactly as in the unfused sequence; on a bracket miss the finish kernel
// runs the exact select over the SCORE slot this kernel has just written.
// =============================================================================================
template <bool ACC, int B>
__device__ __forceinline__ float4 snip_score4(const float* __restrict__ w, const float* const (&g)[B], const float* __restrict__ s, int e) {
    const float4 wv = ld_nc_f4(w + e);
    float4 gv[B];
#pragma unroll
    for (int b = 0; b < B; ++b) gv[b] = ld_nc_f4(g[b] + e);
    float4 r;
    if (ACC) r = ld_f4(s + e);
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const float4 t = make_float4(fabsf(__fmul_rn(wv.x, gv[b].x)), fabsf(__fmul_rn(wv.y, gv[b].y)),
                                     fabsf(__fmul_rn(wv.z, gv[b].z)), fabsf(__fmul_rn(wv.w, gv[b].w)));
        if (b == 0 && !ACC) r = t;
        else { r.x = __fadd_rn(r.x, t.x); r.y = __fadd_rn(r.y, t.y); r.z = __fadd_rn(r.z, t.z); r.w = __fadd_rn(r.w, t.w); }
    }
    return r;
}
template <bool ACC, int B>
__device__ __forceinline__ float snip_score1(const float* __restrict__ w, const float* const (&g)[B], const float* __restrict__ s, int e) {
    float r = ACC ? s[e] : 0.f;
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const float t = fabsf(__fmul_rn(w[e], g[b][e]));
        r = (b == 0 && !ACC) ? t : __fadd_rn(r, t);
    }
    return r;
}

// Sample granules for the fused SNIP path: 8 consecutive keys (one 32-byte DRAM sector per array) instead of the 4 of
// k_select_sample, 4 granules per chunk (1/128 of the keys).  Every granule costs 1 + B sector reads from B + 1
// different arrays, so the kernel is bound by the random-access rate of HBM, not by bytes: the first version
// (16 x 4 keys per chunk) took 33 us for ResNet-50, twice the plain sample.  Neighbouring keys are correlated
// (same filter), so the bracket margin is 12 sigma instead of 8 (SampleArgs::sigmas).
constexpr int kSnipGranulesPerChunk = 4;
// REFRESH: the gradient tensors of this build are new (fresh backward passes): their per-chunk pointer tables are written by
// this kernel — the thread that samples a chunk's first granule derives the chunk's pointers from the segment pointers in the
// kernel arguments anyway — instead of by a launch of their own before it (b200p_snip_mask_build_refresh).
struct RefreshArgs { PtrPackBig pack; const int32_t* chunk_seg; const int64_t* chunk_elem0; int n_seg; };
template <bool ACC, int B, bool REFRESH>
__device__ __forceinline__ void snip_sample_body(const SampleArgs& a, ChunkTab w_tab, const GradTabs& g_tabs, ChunkTab s_tab, const RefreshArgs* rf) {
    __shared__ uint32_t s_hist[kHistBins];
    __shared__ unsigned long long s_warp[9];
    __shared__ uint32_t s_bkt[2];
    for (int b = threadIdx.x; b < kHistBins; b += kThreads) s_hist[b] = 0;
    __syncthreads();
    const int64_t gtid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nthreads = (int64_t)gridDim.x * kThreads;
    const int64_t slots = a.n_chunks * kSnipGranulesPerChunk;
    for (int64_t sl = gtid; sl < slots; sl += nthreads) {
        const int64_t c = sl >> 2;
        const int i = (int)(sl & 3);
        const int e0 = 1024 * i + 8 * (int)((c * 7 + 5 * i) & 127);
        const int n = __ldg(a.chunk_n + c);
        const float* w = chunk_ptr<const float>(w_tab, c);
        const float* sc = ACC ? chunk_ptr<const float>(s_tab, c) : nullptr;
        const float* g[B];
        if (REFRESH) {
            const int seg = __ldg(rf->chunk_seg + c);
            const int64_t el = __ldg(rf->chunk_elem0 + c);
#pragma unroll
            for (int b = 0; b < B; ++b) {
                g[b] = reinterpret_cast<const float*>(rf->pack.p[b * rf->n_seg + seg]) + el;
                if (i == 0) const_cast<void**>(g_tabs.t[b])[c] = const_cast<float*>(g[b]);       // every chunk has a granule 0
            }
        } else {
#pragma unroll
            for (int b = 0; b < B; ++b) g[b] = chunk_ptr<const float>(g_tabs.t[b], c);
        }
        if (e0 >= n) continue;
        float v[8]; int cnt = 0;
        if (a.vec_ok && e0 + 7 < n) {
            const float4 t0 = snip_score4<ACC, B>(w, g, sc, e0), t1 = snip_score4<ACC, B>(w, g, sc, e0 + 4);
            v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w; cnt = 8;
        } else {
            for (int q = 0; q < 8; ++q) v[q] = 0.f;
            for (int q = 0; q < 8 && e0 + q < n; ++q) { v[q] = snip_score1<ACC, B>(w, g, sc, e0 + q); cnt = q + 1; }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q < cnt) atomicAdd(&s_hist[key_of(v[q]) >> 19], 1u);
    }
    __syncthreads();
    flush_hist(s_hist, a.hist, true);
    if (!last_cta_arrives(a.ticket)) return;
    sample_tail(a, nullptr, s_warp, s_bkt, s_hist);
}
template <bool ACC, int B>
__global__ void __launch_bounds__(kThreads)
k_snip_sample(SampleArgs a, ChunkTab w_tab, GradTabs g_tabs, ChunkTab s_tab) {
    snip_sample_body<ACC, B, false>(a, w_tab, g_tabs, s_tab, nullptr);
}
template <bool ACC, int B>
__global__ void __launch_bounds__(kThreads)
k_snip_refresh_sample(SampleArgs a, ChunkTab w_tab, GradTabs g_tabs, ChunkTab s_tab, const __grid_constant__ RefreshArgs rf) {
    snip_sample_body<ACC, B, true>(a, w_tab, g_tabs, s_tab, &rf);
}

template <bool ACC, int B>
__global__ void __launch_bounds__(kThreads, 4)
k_snip_score_sweep(PassArgs a, ChunkTab w_tab, GradTabs g_tabs, int vec_all) {
    __shared__ uint32_t s_hist[kHistBins];
    __shared__ uint32_t sg_key[kThreads / 32][kStage];
    __shared__ uint32_t sg_pos[kThreads / 32][kStage];
    __shared__ int s_wcount[kThreads / 32];
    __shared__ unsigned long long s_warp[9];
    __shared__ unsigned long long s_below;
    SelState* __restrict__ st = a.st;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool classify = st->sample_ok != 0u;           // written by the sample kernel: uniform over the grid
    const uint32_t lo_b = st->lo_bucket, hi_b = st->hi_bucket;
    const uint32_t base = lo_b << 19, span = (hi_b - lo_b + 1) << 19;
    for (int b = tid; b < kHistBins; b += kThreads) s_hist[b] = 0;
    if (tid == 0) s_below = 0;
    if (lane == 0) s_wcount[warp] = 0;
    if (blockIdx.x == 0 && tid == 0) st->prov_ok = (a.prov != nullptr && classify) ? 1u : 0u;
    if (blockIdx.x == 0) sel_stamp(8);
    __syncthreads();

    auto take = [&](uint32_t key, uint32_t pos) {
        atomicAdd(&s_hist[(key - base) >> kFineShift], 1u);
        const int off = atomicAdd(&s_wcount[warp], 1);
        if (off < kStage) { sg_key[warp][off] = key; sg_pos[warp][off] = pos; }
        else {
            const uint32_t gi = atomicAdd(&st->cand_count, 1u);
            if ((long long)gi < a.cand_capacity) { a.cand_key[gi] = key; a.cand_pos[gi] = pos; }
        }
    };
    auto flush = [&](int n) {
        if (n <= 0) return;
        uint32_t gbase = 0;
        if (lane == 0) gbase = atomicAdd(&st->cand_count, (uint32_t)n);
        gbase = __shfl_sync(0xFFFFFFFFu, gbase, 0);
        for (int i = lane; i < n; i += 32)
            if ((long long)(gbase + i) < a.cand_capacity) { a.cand_key[gbase + i] = sg_key[warp][i]; a.cand_pos[gbase + i] = sg_pos[warp][i]; }
        __syncwarp();
        if (lane == 0) s_wcount[warp] = 0;
        __syncwarp();
    };

    unsigned long long below = 0;
    unsigned int nan_cnt = 0;
    for (int64_t c = a.c_begin + blockIdx.x; c < a.c_end; c += gridDim.x) {
        const int n = __ldg(a.chunk_n + c);
        const float* __restrict__ w = chunk_ptr<const float>(w_tab, c);
        float* __restrict__ s = chunk_ptr<float>(a.key_tab, c);
        const float* g[B];
#pragma unroll
        for (int b = 0; b < B; ++b) g[b] = chunk_ptr<const float>(g_tabs.t[b], c);
        const uint32_t pos0 = (uint32_t)(c * kChunk);
        uint32_t* pw = a.prov ? a.prov + (size_t)c * kWordsPerChunk : nullptr;
        if (vec_all && n == kChunk) {
#pragma unroll 2
            for (int j = 0; j < kVecPerThread; ++j) {
                const int e = 4 * (j * kThreads + tid);
                const float4 r = snip_score4<ACC, B>(w, g, s, e);
                st_f4(s + e, r);
                if (classify) {
                    // four keys: sign bits of (key - base) and (key - base - span) shifted into 4-bit masks (key q -> bit 3 - q)
                    const float f[4] = {r.x, r.y, r.z, r.w};
                    uint32_t lt = 0, in = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t d = (__float_as_uint(f[q]) & 0x7FFFFFFFu) - base;
                        lt = __funnelshift_l(d, lt, 1);
                        in = __funnelshift_l(d - span, in, 1);
                    }
                    lt &= 0xFu;
                    below += __popc(lt);
                    if (pw) {
                        // NaN scores are pruned (NaN > thr is false, train.py:316): the patching emit never sees them
                        const uint32_t nanb = (r.x != r.x ? 8u : 0u) | (r.y != r.y ? 4u : 0u) | (r.z != r.z ? 2u : 0u) | (r.w != r.w ? 1u : 0u);
                        lt |= nanb;
                        nan_cnt += __popc(nanb);
                        const uint32_t word = gather_nibbles(__brev(~lt & 0xFu) >> 28);      // bit q = key q at or above the bracket base
                        if ((tid & 7) == 0) pw[vec_word_index(j)] = word;
                    }
                    uint32_t match = in & ~lt & 0xFu;
                    while (match) {                           // divergent, rare (~1-2 % of the keys)
                        const int bit = __ffs(match) - 1;
                        match &= match - 1;
                        const int q = 3 - bit;
                        const float fv = q == 0 ? r.x : q == 1 ? r.y : q == 2 ? r.z : r.w;     // no dynamic indexing: keeps r in registers
                        take(__float_as_uint(fv) & 0x7FFFFFFFu, pos0 + (uint32_t)(e + q));
                    }
                }
            }
        } else {
            for (int it = 0; it < kChunk / kThreads; ++it) {     // warp-uniform trip count: one mask word per warp and step
                const int e = it * kThreads + tid;
                const bool valid = e < n;
                uint32_t k = 0u;
                if (valid) {
                    const float r = snip_score1<ACC, B>(w, g, s, e);
                    s[e] = r;
                    k = key_of(r);
                }
                if (classify) {
                    if (pw) {
                        const uint32_t word = __ballot_sync(0xFFFFFFFFu, valid && k >= base && k != kNanKey);
                        if (lane == 0) pw[e >> 5] = word;
                        if (valid && k == kNanKey) ++nan_cnt;
                    }
                    if (valid) {
                        if (k < base) ++below;
                        else if (k - base < span) take(k, pos0 + (uint32_t)e);
                    }
                }
            }
        }
        if (classify) {
            __syncwarp();
            const int filled = __shfl_sync(0xFFFFFFFFu, s_wcount[warp], 0);       // uniform decision, see sweep_loop
            if (filled >= kStage / 2) flush(filled < kStage ? filled : kStage);
        }
    }
    if (!classify) return;                            // the finish kernel runs the exact select over the scores
    __syncwarp();
    {
        const int filled = __shfl_sync(0xFFFFFFFFu, s_wcount[warp], 0);
        flush(filled < kStage ? filled : kStage);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xFFFFFFFFu, below, o);
    if (lane == 0 && below) atomicAdd(&s_below, below);
    __syncthreads();
    flush_hist(s_hist, a.hist, false);
    if (tid == 0 && s_below) atomicAdd(a.hist + kHistBins + 1, s_below);
    if (nan_cnt) atomicAdd(a.hist + kHistBins + 2, (unsigned long long)nan_cnt);
    if (tid == 0) atomicMax(&g_sel_stamps[9], sel_globaltimer());
    if (!last_cta_arrives(a.ticket)) return;
    bracket_tail(a, nullptr, base, span, s_warp, s_hist);
}

// ---- B: finish (cooperative launch) ----------------------------------------------------------------
__device__ __forceinline__ void grid_barrier() { cooperative_groups::this_grid().sync(); }

// EMIT: the kernel also emits the mask (b200p_mask_build: the sweep wrote the provisional mask straight into the
// destination): once a CTA knows the exact key it patches its share of the candidates' bits — every CTA derives the key
// itself from the global window histogram, so there is no barrier between "key known" and "mask patched", and the
// separate emit launch is gone.
template <bool EMIT>
__global__ void __launch_bounds__(kThreads, 3)
k_select_finish(PassArgs a, uint32_t* __restrict__ chunk_ties, int64_t n_chunks, EmitArgs em) {
    __shared__ uint32_t s_hist[kHistBins];
    __shared__ unsigned long long s_part[kThreads];
    __shared__ unsigned long long s_warp[9];
    __shared__ int s_last;
    SelState* __restrict__ st = a.st;
    const bool exact = st->miss != 0u;               // written by the previous kernel: uniform over the grid
    const bool prov_ok = st->prov_ok != 0u;          // idem
    PatchVals pv;
    bool pv_valid = false, use_list = false;
    if (exact) {
        if (blockIdx.x == 0 && threadIdx.x == 0) init_state(st, a.k, a.mode, a.allow_collect, 1u);
        grid_barrier();
        pass_body<0>(a, s_hist); grid_barrier();
        if (blockIdx.x == 0) scan_and_advance(0, a.hist, st, a.cand_capacity);
        grid_barrier();
        pass_body<1>(a, s_hist); grid_barrier();
        if (blockIdx.x == 0) scan_and_advance(1, a.hist, st, a.cand_capacity);
        grid_barrier();
        pass_body<2>(a, s_hist); grid_barrier();
        if (blockIdx.x == 0) {
            scan_and_advance(2, a.hist, st, a.cand_capacity);
            if (threadIdx.x == 0) st->passes_full = st->collect ? 3u : 4u;
        }
        grid_barrier();
        // ties (EXACT_K with quota < n_equal): per-chunk counts, then the ordered scan
        if (a.mode == B200P_MODE_EXACT_K && st->need_ties) {
            tie_count_body(a.chunk_n, a.key_tab, a.old_mask, st, a.cand_key, a.cand_pos, chunk_ties, 0, n_chunks, a.vec_ok);
            grid_barrier();
            if (blockIdx.x == 0) tie_scan_body(st, chunk_ties, 0, n_chunks, 0ull, nullptr, 0, s_part);
            if (EMIT) grid_barrier();
        }
    } else {
        // window pass over the candidates: exact key among the 1024 keys of the window.  Everything a CTA needs from the
        // state is read BEFORE the barrier: afterwards CTA 0 rewrites the state while the others are still deriving the key.
        const uint32_t n = st->cand_count, win_lo = st->win_lo, mode = st->mode, n_nan = st->pad_[1];
        const unsigned long long k = st->k, n_less0 = st->n_less, n_valid = st->n_valid;
        for (int b = threadIdx.x; b < kWindow; b += kThreads) s_hist[b] = 0;
        __syncthreads();
        const uint32_t stride = gridDim.x * kThreads;
        for (uint32_t i0 = blockIdx.x * kThreads + threadIdx.x; i0 < n; i0 += 8 * stride) {
            uint32_t kk[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) kk[u] = i0 + u * stride < n ? __ldg(a.cand_key + i0 + u * stride) : 0xFFFFFFFFu;   // independent loads
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t d = kk[u] - win_lo;
                if (i0 + u * stride < n && d < (uint32_t)kWindow) atomicAdd(&s_hist[d], 1u);
            }
        }
        __syncthreads();
        for (int b = threadIdx.x; b < kWindow; b += kThreads) {
            const uint32_t v = s_hist[b];
            if (v) atomicAdd(a.hist + b, (unsigned long long)v);
        }
        grid_barrier();
        // every CTA: prefix over the 1024 global counts -> the key that holds rank k
        unsigned long long local[kBinsPerThread];
#pragma unroll
        for (int i = 0; i < kBinsPerThread; ++i) {
            const int b = threadIdx.x * kBinsPerThread + i;
            local[i] = b < kWindow ? __ldcg(a.hist + b) : 0ull;
        }
        unsigned long long total;
        unsigned long long running = block_prefix16(local, s_warp, total);
        __shared__ unsigned long long s_pick[3];      // bin, count before it, count in it
        if (threadIdx.x == 0) s_pick[0] = ~0ull;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kBinsPerThread; ++i) {
            const unsigned long long v = local[i];
            if (v != 0 && running < k && k <= running + v) { s_pick[0] = threadIdx.x * kBinsPerThread + i; s_pick[1] = running; s_pick[2] = v; }
            running += v;
        }
        __syncthreads();
        // the last CTA to have read the histogram clears it for the next select (nobody waits for that)
        if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1 ? 1 : 0;
        __syncthreads();
        if (s_last) { clear_hist(a.hist); if (threadIdx.x == 0) *a.ticket = 0u; }
        const bool found = s_pick[0] != ~0ull;      // always, after a verified bracket
        const uint32_t key = win_lo + (uint32_t)s_pick[0];
        const unsigned long long before = s_pick[1], v = s_pick[2];
        const uint32_t need_ties = (found && mode == B200P_MODE_EXACT_K && (k - before) < v) ? 1u : 0u;
        if (blockIdx.x == 0 && threadIdx.x == 0 && found) {
            st->n_less = n_less0 + before;
            st->k = k - before;
            st->prefix = key; st->thr_key = key; st->threshold = key_to_float(key);
            st->n_equal = v; st->quota = k - before;
            st->need_ties = need_ties;
            st->tie_chunk = -1; st->tie_resid = 0; st->tie_seen = 0;
        }
        if (need_ties && v <= (unsigned long long)kTieListCap && EMIT && prov_ok) {
            // few ties and the mask is patched right here: the patch pass gathers them, CTA 0 prunes them by position afterwards
            pv.cand_count = n; pv.thr_key = key; pv.need_ties = 0u; pv.tie_resid = 0u; pv.tie_chunk = -1;
            pv.n_kept = n_valid - (n_less0 + before) - (k - before);
            pv_valid = true; use_list = true;
        } else if (need_ties && v <= (unsigned long long)kTieListCap) {
            tie_list_gather(key, n, a.cand_key, a.cand_pos, chunk_ties + n_chunks);
            grid_barrier();
            if (blockIdx.x == 0) tie_list_pick(st, chunk_ties + n_chunks, 0, n_chunks, 0ull, nullptr, 0, reinterpret_cast<uint32_t*>(s_part));
            if (EMIT) grid_barrier();
        } else if (need_ties) {
            tie_count_vals(a.chunk_n, a.key_tab, a.old_mask, 1u, key, 1u, n, a.cand_key, a.cand_pos, chunk_ties, 0, n_chunks, a.vec_ok);
            grid_barrier();
            if (blockIdx.x == 0) tie_scan_body(st, chunk_ties, 0, n_chunks, 0ull, nullptr, 0, s_part);
            if (EMIT) grid_barrier();
        } else if (found) {
            pv.cand_count = n; pv.thr_key = key; pv.need_ties = 0u; pv.tie_resid = 0u; pv.tie_chunk = -1;
            pv.n_kept = n_valid - (n_less0 + before) - (mode == B200P_MODE_SNIP_STRICT ? v + (key != kNanKey ? n_nan : 0u) : (k - before));
            pv_valid = true;
        }
    }
    if (!EMIT) return;
    if (prov_ok && !exact) {
        if (!pv_valid) pv = patch_vals_from_state(st, em.mode);          // after the tie barriers: the state is final
        emit_patch_body(em.chunk_n, em.key_tab, em.old_mask, st, em.cand_key, em.cand_pos, em.prov, em.mode, em.n_chunks, pv,
                        use_list ? chunk_ties + n_chunks : nullptr);
        if (use_list) {
            grid_barrier();
            if (blockIdx.x == 0) tie_list_pick(st, chunk_ties + n_chunks, 0, n_chunks, 0ull, nullptr, 0, reinterpret_cast<uint32_t*>(s_part), em.prov);
        }
    } else {
        // no candidate list to patch (exact select, histogram mode): full pass over the keys; the state is final here
        if (!exact || !(a.mode == B200P_MODE_EXACT_K && st->need_ties)) grid_barrier();
        emit_full_body(em, 0, n_chunks);
    }
}

// ---- sharded tail: finish + ties + emit + mask push in ONE cooperative launch ----------------------------------------------
// The stages after the sweep are tiny and strictly ordered; as five launches (finish, tie count, tie scan, emit, push)
// they cost ~60 us of launch latency and tails on a slice that streams in 20.  Here grid-wide barriers separate them:
//   window histogram -> [CTA 0: all-gather over the ranks, exact key] -> (ties: count -> [CTA 0: ordered scan]) ->
//   patch the own candidates' bits -> push the own mask words into every window -> [CTA 0: tell the peers, wait for theirs]
__global__ void __launch_bounds__(kThreads, 3)
k_sharded_tail(PassArgs a, CommDev comm, uint32_t seq_gather, uint32_t seq_mask, unsigned long long* __restrict__ rank_ties,
               uint32_t* __restrict__ chunk_ties, EmitArgs em, int64_t c0, int64_t c1) {
    __shared__ uint32_t s_hist[kWindow];
    __shared__ unsigned long long s_part[kThreads];
    __shared__ unsigned long long s_warp[9];
    __shared__ int s_bin;
    SelState* __restrict__ st = a.st;
    // measurement aid (b200p_comm_trace, channel "barrier"): [0] start [1] window histogram flushed [2] exact key known
    // [3] ties resolved [4] own bits patched [5] own words pushed [6] every rank's words here
    unsigned long long* tr = comm_trace_slot(comm, CH_BARRIER, 0u);
    auto stamp = [&](int i) { if (blockIdx.x == 0 && threadIdx.x == 0) tr[i] = comm_globaltimer(); };
    stamp(0);
    if (st->miss) return;                            // same on every rank (identical state): the builder reruns the staged exact select
    for (int b = threadIdx.x; b < kWindow; b += kThreads) s_hist[b] = 0;
    __syncthreads();
    {
        const uint32_t n = st->cand_count, win_lo = st->win_lo;
        const uint32_t stride = gridDim.x * kThreads;
        for (uint32_t i0 = blockIdx.x * kThreads + threadIdx.x; i0 < n; i0 += 8 * stride) {
            uint32_t kk[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) kk[u] = i0 + u * stride < n ? __ldg(a.cand_key + i0 + u * stride) : 0xFFFFFFFFu;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t d = kk[u] - win_lo;
                if (i0 + u * stride < n && d < (uint32_t)kWindow) atomicAdd(&s_hist[d], 1u);
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kWindow; b += kThreads) {
        const uint32_t v = s_hist[b];
        if (v) atomicAdd(a.hist + b, (unsigned long long)v);
    }
    grid_barrier();
    stamp(1);
    if (blockIdx.x == 0) sharded_gather_pick(a, comm, seq_gather, rank_ties, s_warp, &s_bin, s_hist);
    grid_barrier();
    stamp(2);
    if (st->miss) return;                            // uniform after the barrier
    // few ties (the normal case) and a candidate list to patch: the patch pass gathers the tied candidates, CTA 0 prunes the
    // right ones by position afterwards — one pass over the candidates and one barrier less than resolving the ties first
    const bool list_mode = a.mode == B200P_MODE_EXACT_K && st->need_ties && st->collect && st->prov_ok &&
                           st->n_equal <= (unsigned long long)kTieListCap;         // uniform: the state is final since the barrier
    if (a.mode == B200P_MODE_EXACT_K && st->need_ties && !list_mode) {
        tie_count_body(a.chunk_n, a.key_tab, a.old_mask, st, a.cand_key, a.cand_pos, chunk_ties, c0, c1, a.vec_ok);
        grid_barrier();
        if (blockIdx.x == 0) tie_scan_body(st, chunk_ties, c0, c1, 0ull, rank_ties, comm.rank, s_part);
        grid_barrier();
    }
    stamp(3);
    if (st->prov_ok) {
        PatchVals pv = patch_vals_from_state(st, em.mode);
        emit_patch_body(em.chunk_n, em.key_tab, em.old_mask, st, em.cand_key, em.cand_pos, em.prov, em.mode, em.n_chunks, pv,
                        list_mode ? chunk_ties + em.n_chunks : nullptr);
        if (list_mode) {
            grid_barrier();
            if (blockIdx.x == 0) tie_list_pick(st, chunk_ties + em.n_chunks, c0, c1, 0ull, rank_ties, comm.rank, reinterpret_cast<uint32_t*>(s_part), em.prov);
        }
    } else emit_full_body(em, c0, c1);
    grid_barrier();                                  // every bit of the own slice is final
    stamp(4);
    comm_push_mask_words(comm, c0 * kWordsPerChunk, c1 * kWordsPerChunk);
    grid_barrier();                                  // every CTA's remote stores are issued and fenced
    stamp(5);
    if (blockIdx.x == 0) comm_signal_and_wait(comm, CH_MASK, seq_mask);
    stamp(6);
}

static unsigned int* ticket_ptr(b200p_plan* p);
static int coop_ctas(b200p_plan* p) {
    if (p->coop_ctas_per_sm == 0) {
        int coop = 0, occ = 0, occ2 = 0;
        B200P_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device));
        if (coop) {
            B200P_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_select_finish<false>, kThreads, 0));
            B200P_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_select_finish<true>, kThreads, 0));
            if (occ2 < occ) occ = occ2;
        }
        p->coop_ctas_per_sm = occ > 0 ? (occ > 3 ? 3 : occ) : -1;
    }
    return B200P_OK;
}
// B: finish (cooperative: grid-wide barriers between its phases).  With p->fuse_emit the kernel also emits the mask.
static int launch_finish(b200p_plan* p, PassArgs& a, int key_source, int mode, const uint32_t* d_old_mask, cudaStream_t st) {
    int64_t work = p->n_chunks;
    const int64_t cblocks = (p->cand_capacity + 8 * kThreads - 1) / (8 * kThreads);
    if (cblocks > work) work = cblocks;
    // one CTA per SM: the phases between the grid-wide barriers are tiny, the barriers are not
    int grid = p->num_sms;
    if (work < grid) grid = (int)(work < 1 ? 1 : work);
    uint32_t* ties = p->d_chunk_ties; int64_t n_chunks = p->n_chunks;
    EmitArgs em;
    uint32_t* prov = p->prov_target ? p->prov_target : p->d_prov;
    fill_patch_emit_args(p, em, key_source, mode, d_old_mask, prov, prov);
    void* args[] = {(void*)&a, (void*)&ties, (void*)&n_chunks, (void*)&em};
    const bool fuse = p->fuse_emit && p->prov_target != nullptr;
    B200P_CUDA(cudaLaunchCooperativeKernel(fuse ? (const void*)k_select_finish<true> : (const void*)k_select_finish<false>, dim3(grid), dim3(kThreads),
                                           args, 0, st));
    p->emit_done = fuse;
    return B200P_OK;
}

static unsigned int* ticket_ptr(b200p_plan* p) {
    // the ticket counter lives in the padding at the end of the state block
    return (unsigned int*)((char*)p->d_state + offsetof(SelState, pad_));
}

// plain launch returning the error (the sample / sweep kernels; a cluster launch lived here for a while, see flush_hist)
template <typename K, typename... Args>
static cudaError_t launch_kernel(K kern, int grid, cudaStream_t st, Args... args) {
    kern<<<grid, kThreads, 0, st>>>(args...);
    return cudaGetLastError();
}

}  // namespace b200p

using namespace b200p;

// B200P_OPT_REUSE_SAMPLE: the caller promises that the keys (and the old mask) are unchanged since the select that
// filled the cache — true inside a sparsity sweep over fixed weights (BASELINE config 5); re-binding the key slot or
// any other key source / old mask / chunk range drops the cache.
// The bracket is a whole number of 12-bit buckets (each holds a few % of the keys near the mode), so its width stops
// shrinking once the sample is a few hundred thousand float4: the 16 per chunk that suit ResNet-50 (100 k float4) are
// 892 k random 32-byte sector reads on ViT-L/16 — 56 us, a quarter of the sweep.  Large sets are sampled more thinly.
static int sample_slot_shift(int64_t n_chunks) {
    int shift = 4;                                                     // kSampleSlotsPerChunk = 16
    while (shift > 1 && (n_chunks << shift) > 480 * 1024) --shift;
    return shift;
}
static bool sample_cache_hit(b200p_plan* p, int key_source, const uint32_t* d_old_mask, int64_t c0, int64_t c1) {
    return p->reuse_sample && p->sample_cache_valid && p->sample_cache_key == key_source && p->sample_cache_mask == d_old_mask &&
           p->sample_cache_c0 == c0 && p->sample_cache_c1 == c1 && p->sample_cache_tab == (const void*)p->d_tab[key_source == B200P_KEY_ABS_W ? B200P_SLOT_W : B200P_SLOT_SCORE];
}
static unsigned long long* sample_cache_arm(b200p_plan* p, int key_source, const uint32_t* d_old_mask, int64_t c0, int64_t c1) {
    if (!p->reuse_sample) { p->sample_cache_valid = false; return nullptr; }
    p->sample_cache_valid = true; p->sample_cache_key = key_source; p->sample_cache_mask = d_old_mask; p->sample_cache_c0 = c0; p->sample_cache_c1 = c1;
    p->sample_cache_tab = (const void*)p->d_tab[key_source == B200P_KEY_ABS_W ? B200P_SLOT_W : B200P_SLOT_SCORE];
    return p->d_sample_cache;
}

static int check_key_source(b200p_plan* p, int key_source, const char* who) {
    const int slot = key_source == B200P_KEY_ABS_W ? B200P_SLOT_W : B200P_SLOT_SCORE;
    if (key_source != B200P_KEY_ABS_W && key_source != B200P_KEY_SCORE) {
        set_error(std::string(who) + ": bad key_source"); return B200P_EINVAL;
    }
    if (!p->bound[slot]) { set_error(std::string(who) + ": key slot is not bound"); return B200P_ESTATE; }
    return B200P_OK;
}
static inline int key_slot(int key_source) { return key_source == B200P_KEY_ABS_W ? B200P_SLOT_W : B200P_SLOT_SCORE; }

extern "C" int b200p_select_begin(b200p_plan* p, uint64_t k, int mode, int allow_collect, void* stream) {
    B200P_REQUIRE(p != nullptr, B200P_EINVAL, "select_begin: null plan");
    B200P_REQUIRE(mode == B200P_MODE_SNIP_STRICT || mode == B200P_MODE_EXACT_K, B200P_EINVAL, "select_begin: bad mode");
    B200P_CUDA(cudaSetDevice(p->device));
    p->prov_armed = false;
    k_select_init<<<1, 256, 0, (cudaStream_t)stream>>>(p->d_state, p->d_hist, ticket_ptr(p), k, (uint32_t)mode,
                                                      allow_collect ? 1u : 0u);
    B200P_LAUNCH_CHECK("k_select_init");
    return B200P_OK;
}

// grid of a pass: enough CTAs for the chunk range and, for pass 2 / ties, for the candidate buffer
static int pass_grid(b200p_plan* p, int64_t chunks, bool cand_too, int ctas_per_sm = 3) {
    int64_t work = chunks;
    if (cand_too) {
        const int64_t blocks = (p->cand_capacity + 8 * kThreads - 1) / (8 * kThreads);
        if (blocks > work) work = blocks;
    }
    return p->grid_for(work, ctas_per_sm);
}

static void fill_pass_args(b200p_plan* p, PassArgs& a, int key_source, const uint32_t* d_old_mask, int64_t c0, int64_t c1,
                           int fuse_scan, int fuse_init, uint64_t k, int mode, int allow_collect, bool with_prov = false);

static int launch_pass(b200p_plan* p, int pass, int key_source, const uint32_t* d_old_mask,
                       int64_t c0, int64_t c1, int fuse_scan, int fuse_init, uint64_t k, int mode, int allow_collect,
                       cudaStream_t st, bool with_prov = false) {
    PassArgs a;
    fill_pass_args(p, a, key_source, d_old_mask, c0, c1, fuse_scan, fuse_init, k, mode, allow_collect, with_prov);
    const int grid = pass_grid(p, c1 - c0, pass == 2, pass == 0 ? 4 : 3);
    switch (pass) {
        case 0: k_select_pass<0><<<grid, kThreads, 0, st>>>(a); break;
        case 1: k_select_pass<1><<<grid, kThreads, 0, st>>>(a); break;
        default: k_select_pass<2><<<grid, kThreads, 0, st>>>(a); break;
    }
    B200P_LAUNCH_CHECK("k_select_pass");
    return B200P_OK;
}

extern "C" int b200p_select_hist(b200p_plan* p, int pass, int key_source, const uint32_t* d_old_mask,
                                 int64_t chunk_begin, int64_t chunk_end, void* stream) {
    B200P_REQUIRE(p != nullptr, B200P_EINVAL, "select_hist: null plan");
    B200P_REQUIRE(pass >= 0 && pass <= 2, B200P_EINVAL, "select_hist: pass must be 0..2");
    int rc = check_key_source(p, key_source, "select_hist"); if (rc) return rc;
    if (chunk_end < 0) chunk_end = p->n_chunks;
    B200P_REQUIRE(chunk_begin >= 0 && chunk_begin <= chunk_end && chunk_end <= p->n_chunks, B200P_EINVAL, "select_hist: bad chunk range");
    B200P_CUDA(cudaSetDevice(p->device));
    // pass 2 always launches: in collect mode it histograms this rank's candidate buffer
    if (chunk_end > chunk_begin || pass == 2)
        return launch_pass(p, pass, key_source, d_old_mask, chunk_begin, chunk_end, 0, 0, 0, 0, 0, (cudaStream_t)stream);
    return B200P_OK;
}

extern "C" int b200p_select_scan(b200p_plan* p, int pass, void* stream) {
    B200P_REQUIRE(p != nullptr, B200P_EINVAL, "select_scan: null plan");
    B200P_REQUIRE(pass >= 0 && pass <= 2, B200P_EINVAL, "select_scan: pass must be 0..2");
    B200P_CUDA(cudaSetDevice(p->device));
    k_select_scan<<<1, kScanThreads, 0, (cudaStream_t)stream>>>(pass, p->d_hist, p->d_state, p->cand_capacity);
    B200P_LAUNCH_CHECK("k_select_scan");
    return B200P_OK;
}

static int launch_tie_count(b200p_plan* p, int key_source, const uint32_t* d_old_mask,
                            int64_t chunk_begin, int64_t chunk_end, cudaStream_t st) {
    const int slot = key_slot(key_source);
    k_tie_count<<<pass_grid(p, chunk_end - chunk_begin, true), kThreads, 0, st>>>(p->d_chunk_n, p->tab(slot), d_old_mask,
        p->d_state, p->d_cand_key, p->d_cand_pos, p->d_chunk_ties, chunk_begin, chunk_end, p->vec_ok[slot] ? 1 : 0);
    B200P_LAUNCH_CHECK("k_tie_count");
    return B200P_OK;
}

static int check_tie_args(b200p_plan* p, int key_source, int64_t& chunk_begin, int64_t& chunk_end, const char* who) {
    if (!p) { set_error(std::string(who) + ": null plan"); return B200P_EINVAL; }
    int rc = check_key_source(p, key_source, who); if (rc) return rc;
    if (chunk_end < 0) chunk_end = p->n_chunks;
    if (!(chunk_begin >= 0 && chunk_begin <= chunk_end && chunk_end <= p->n_chunks)) {
        set_error(std::string(who) + ": bad chunk range"); return B200P_EINVAL;
    }
    return B200P_OK;
}

extern "C" int b200p_select_ties(b200p_plan* p, int key_source, const uint32_t* d_old_mask,
                                 int64_t chunk_begin, int64_t chunk_end, uint64_t tie_offset, void* stream) {
    int rc = check_tie_args(p, key_source, chunk_begin, chunk_end, "select_ties"); if (rc) return rc;
    B200P_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_tie_count(p, key_source, d_old_mask, chunk_begin, chunk_end, st); if (rc) return rc;
    k_tie_scan<<<1, 1024, 0, st>>>(p->d_state, p->d_chunk_ties, chunk_begin, chunk_end, (unsigned long long)tie_offset, nullptr, 0);
    B200P_LAUNCH_CHECK("k_tie_scan");
    return B200P_OK;
}

extern "C" int b200p_select_ties_count(b200p_plan* p, int key_source, const uint32_t* d_old_mask,
                                       int64_t chunk_begin, int64_t chunk_end, uint64_t* d_local_count, void* stream) {
    int rc = check_tie_args(p, key_source, chunk_begin, chunk_end, "select_ties_count"); if (rc) return rc;
    B200P_REQUIRE(d_local_count != nullptr, B200P_EINVAL, "select_ties_count: null output");
    B200P_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_tie_count(p, key_source, d_old_mask, chunk_begin, chunk_end, st); if (rc) return rc;
    k_tie_total<<<1, 1024, 0, st>>>(p->d_state, p->d_chunk_ties, chunk_begin, chunk_end, (unsigned long long*)d_local_count);
    B200P_LAUNCH_CHECK("k_tie_total");
    return B200P_OK;
}

extern "C" int b200p_select_ties_scan(b200p_plan* p, int64_t chunk_begin, int64_t chunk_end,
                                      const uint64_t* d_counts, int n_before, void* stream) {
    B200P_REQUIRE(p != nullptr, B200P_EINVAL, "select_ties_scan: null plan");
    if (chunk_end < 0) chunk_end = p->n_chunks;
    B200P_REQUIRE(chunk_begin >= 0 && chunk_begin <= chunk_end && chunk_end <= p->n_chunks, B200P_EINVAL, "select_ties_scan: bad chunk range");
    B200P_REQUIRE(n_before == 0 || d_counts != nullptr, B200P_EINVAL, "select_ties_scan: null count table");
    B200P_CUDA(cudaSetDevice(p->device));
    k_tie_scan<<<1, 1024, 0, (cudaStream_t)stream>>>(p->d_state, p->d_chunk_ties, chunk_begin, chunk_end, 0ull,
                                                       (const unsigned long long*)d_counts, n_before);
    B200P_LAUNCH_CHECK("k_tie_scan");
    return B200P_OK;
}

static void fill_pass_args(b200p_plan* p, PassArgs& a, int key_source, const uint32_t* d_old_mask, int64_t c0, int64_t c1,
                           int fuse_scan, int fuse_init, uint64_t k, int mode, int allow_collect, bool with_prov) {
    const int slot = key_slot(key_source);
    a.prov = with_prov ? (p->prov_target ? p->prov_target : p->d_prov) : nullptr;
    a.chunk_n = p->d_chunk_n; a.key_tab = p->tab(slot); a.old_mask = d_old_mask; a.hist = p->d_hist; a.st = p->d_state;
    a.cand_key = p->d_cand_key; a.cand_pos = p->d_cand_pos; a.ticket = ticket_ptr(p); a.cand_capacity = p->cand_capacity;
    a.c_begin = c0; a.c_end = c1; a.vec_ok = p->vec_ok[slot] ? 1 : 0; a.fuse_scan = fuse_scan;
    a.fuse_init = fuse_init; a.k = k; a.mode = (uint32_t)mode; a.allow_collect = allow_collect ? 1u : 0u;
    a.comm_seq = 0u;
}

static int select_kth_exact(b200p_plan* p, int key_source, const uint32_t* d_old_mask, uint64_t k, int mode, cudaStream_t st) {
    // Three launches, no host round trip: the state reset rides in pass 0's last CTA (the global
    // histogram and the ticket are left zeroed by every completed scan), every pass ends with the
    // last-CTA scan.  Pass 2 runs on the candidates gathered by pass 1 unless the bucket overflowed.
    int rc = launch_pass(p, 0, key_source, d_old_mask, 0, p->n_chunks, 1, 1, k, mode, 1, st); if (rc) return rc;
    rc = launch_pass(p, 1, key_source, d_old_mask, 0, p->n_chunks, 1, 0, 0, 0, 0, st, true); if (rc) return rc;
    rc = launch_pass(p, 2, key_source, d_old_mask, 0, p->n_chunks, 1, 0, 0, 0, 0, st); if (rc) return rc;
    if (mode == B200P_MODE_EXACT_K) { rc = b200p_select_ties(p, key_source, d_old_mask, 0, p->n_chunks, 0, (void*)st); if (rc) return rc; }
    return B200P_OK;
}

static int select_kth_sampled(b200p_plan* p, int key_source, const uint32_t* d_old_mask, uint64_t k, int mode, cudaStream_t st) {
    const int slot = key_slot(key_source);
    { int rc = coop_ctas(p); if (rc) return rc; }
    if (p->coop_ctas_per_sm < 0) return select_kth_exact(p, key_source, d_old_mask, k, mode, st);   // no cooperative launch
    // S: 1/16 sample
    SampleArgs sa;
    sa.chunk_n = p->d_chunk_n; sa.key_tab = p->tab(slot); sa.old_mask = d_old_mask; sa.hist = p->d_hist; sa.st = p->d_state;
    sa.ticket = ticket_ptr(p); sa.n_chunks = p->n_chunks; sa.c_begin = 0; sa.comm_seq = 0u; sa.cache = nullptr; sa.vec_ok = p->vec_ok[slot] ? 1 : 0;
    sa.k = k; sa.n_total = (unsigned long long)p->total; sa.mode = (uint32_t)mode; sa.sigmas = 8u;
    sa.slot_shift = sample_slot_shift(p->n_chunks);
    const int64_t sblocks = ((p->n_chunks << sa.slot_shift) + 4 * kThreads - 1) / (4 * kThreads);
    if (sample_cache_hit(p, key_source, d_old_mask, 0, p->n_chunks)) {
        k_sample_from_cache<<<1, kThreads, 0, st>>>(sa, p->d_sample_cache);
        B200P_LAUNCH_CHECK("k_sample_from_cache");
    } else {
        sa.cache = sample_cache_arm(p, key_source, d_old_mask, 0, p->n_chunks);
        B200P_CUDA(launch_kernel(k_select_sample, p->grid_for(sblocks, 1), st, sa, CommDev()));
        B200P_LAUNCH_CHECK("k_select_sample");
    }
    // A: bracket sweep
    PassArgs a;
    fill_pass_args(p, a, key_source, d_old_mask, 0, p->n_chunks, 1, 0, k, mode, 1, true);
    B200P_CUDA(launch_kernel(k_select_bracket, p->grid_for((p->n_chunks + 1) / 2, 4), st, a, CommDev()));
    B200P_LAUNCH_CHECK("k_select_bracket");
    return launch_finish(p, a, key_source, mode, d_old_mask, st);
}

extern "C" int b200p_select_kth(b200p_plan* p, int key_source, const uint32_t* d_old_mask,
                                uint64_t k, int mode, void* stream) {
    B200P_REQUIRE(p != nullptr, B200P_EINVAL, "select_kth: null plan");
    int rc = check_key_source(p, key_source, "select_kth"); if (rc) return rc;
    B200P_REQUIRE(k >= 1 && k <= (uint64_t)p->total, B200P_EINVAL, "select_kth: k must be in [1, N]");
    B200P_REQUIRE(mode == B200P_MODE_SNIP_STRICT || mode == B200P_MODE_EXACT_K, B200P_EINVAL, "select_kth: bad mode");
    B200P_CUDA(cudaSetDevice(p->device));
    // the emit that follows with the same arguments can patch the provisional mask instead of re-reading the keys
    p->prov_armed = true; p->prov_key_source = key_source; p->prov_mode = mode; p->prov_old_mask = d_old_mask;
    p->prov_c0 = 0; p->prov_c1 = p->n_chunks; p->emit_done = false;
    if (p->select_impl == B200P_SELECT_EXACT) return select_kth_exact(p, key_source, d_old_mask, k, mode, (cudaStream_t)stream);
    return select_kth_sampled(p, key_source, d_old_mask, k, mode, (cudaStream_t)stream);
}

// ---- parameter-sharded select + emit + mask all-gather over peer memory ----------------------------------
// Rank r owns chunks [c0, c1).  Sequence (all on `stream`, no host round trip, no NCCL):
//   S  k_select_sample   over the own range; last CTA: all-reduce of the sample histogram + alive count -> bracket
//   A  k_select_bracket  ONE sweep of the own range; provisional mask words go straight into the window's mask area;
//                        last CTA: all-reduce of {fine histogram, below} -> verified bracket, 1024-key window
//   B  k_sharded_finish  window histogram of the own candidates; last CTA: all-gather -> exact key, tie bookkeeping
//   T  k_tie_count / k_tie_scan (EXACT_K): which of the own tied keys are pruned (ties in lower ranks come first)
//   E  k_emit_masks      patches the own candidates' bits
//   P  k_mask_push       own words -> every window; the kernel ends when every rank's words have arrived here
// A bracket miss (never observed on real weight sets; constant tensors do it) leaves SelState::miss = 1 on every rank:
// the host side checks it where it reads the result anyway and reruns the staged exact select.
extern "C" int b200p_sharded_mask_build(b200p_plan* p, b200p_comm* c, int key_source, const uint32_t* d_old_mask, uint64_t k, int mode,
                                        int64_t chunk_begin, int64_t chunk_end, int stages, void* stream) {
    B200P_REQUIRE(p != nullptr && c != nullptr && c->connected, B200P_ESTATE, "sharded_mask_build: null argument / comm not connected");
    int rc = check_key_source(p, key_source, "sharded_mask_build"); if (rc) return rc;
    B200P_REQUIRE(k >= 1 && k <= (uint64_t)p->total, B200P_EINVAL, "sharded_mask_build: k must be in [1, N]");
    B200P_REQUIRE(mode == B200P_MODE_SNIP_STRICT || mode == B200P_MODE_EXACT_K, B200P_EINVAL, "sharded_mask_build: bad mode");
    B200P_REQUIRE(chunk_begin >= 0 && chunk_begin <= chunk_end && chunk_end <= p->n_chunks, B200P_EINVAL, "sharded_mask_build: bad chunk range");
    B200P_REQUIRE(c->mask_words == p->n_chunks * kWordsPerChunk, B200P_EINVAL, "sharded_mask_build: the window's mask area does not match the plan");
    B200P_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int slot = key_slot(key_source);
    const int64_t nc = chunk_end - chunk_begin;
    uint32_t* d_mask = reinterpret_cast<uint32_t*>(c->window + c->lay.mask);
    const CommDev cd = c->dev();
    if (stages == 0) stages = B200P_SHARD_ALL;
    // S
    if (stages & B200P_SHARD_SAMPLE) {
    SampleArgs sa;
    sa.chunk_n = p->d_chunk_n; sa.key_tab = p->tab(slot); sa.old_mask = d_old_mask; sa.hist = p->d_hist; sa.st = p->d_state;
    sa.ticket = ticket_ptr(p); sa.n_chunks = nc; sa.c_begin = chunk_begin; sa.vec_ok = p->vec_ok[slot] ? 1 : 0;
    sa.k = k; sa.n_total = (unsigned long long)p->total; sa.mode = (uint32_t)mode; sa.sigmas = 8u;
    sa.comm_seq = 0u; sa.cache = nullptr;
    sa.slot_shift = sample_slot_shift(p->n_chunks);                   // from the whole set: every rank samples at the same density
    const int64_t sblocks = ((nc << sa.slot_shift) + 4 * kThreads - 1) / (4 * kThreads);
    if (sample_cache_hit(p, key_source, d_old_mask, chunk_begin, chunk_end)) {      // the cache holds the all-reduced histogram: same decision on every rank
        k_sample_from_cache<<<1, kThreads, 0, st>>>(sa, p->d_sample_cache);
        B200P_LAUNCH_CHECK("k_sample_from_cache");
    } else {
        sa.comm_seq = ++c->seq[CH_HIST];
        sa.cache = sample_cache_arm(p, key_source, d_old_mask, chunk_begin, chunk_end);
        B200P_CUDA(launch_kernel(k_select_sample, p->grid_for(sblocks, 1), st, sa, cd));
        B200P_LAUNCH_CHECK("k_select_sample");
    }
    }
    // A
    p->prov_target = d_mask;
    p->prov_armed = true; p->prov_key_source = key_source; p->prov_mode = mode; p->prov_old_mask = d_old_mask;
    p->prov_c0 = chunk_begin; p->prov_c1 = chunk_end;
    PassArgs a;
    fill_pass_args(p, a, key_source, d_old_mask, chunk_begin, chunk_end, 1, 0, k, mode, 1, true);
    if (stages & B200P_SHARD_SWEEP) {
        a.comm_seq = ++c->seq[CH_HIST];
        B200P_CUDA(launch_kernel(k_select_bracket, p->grid_for(nc > 1 ? (nc + 1) / 2 : 1, 4), st, a, cd));
        B200P_LAUNCH_CHECK("k_select_bracket");
    }
    // B + T + E + P in one cooperative launch when the whole tail is wanted
    const int tail_bits = B200P_SHARD_FINISH | B200P_SHARD_TIES | B200P_SHARD_EMIT | B200P_SHARD_PUSH;
    if ((stages & tail_bits) == tail_bits && !getenv("B200P_SHARD_SPLIT_TAIL")) {
        rc = coop_ctas(p); if (rc) return rc;
        if (p->coop_ctas_per_sm > 0) {
            EmitArgs em;
            fill_patch_emit_args(p, em, key_source, mode, d_old_mask, d_mask, d_mask);
            int grid = p->num_sms;
            if (p->coop_grid_limit > 0 && grid > p->coop_grid_limit) grid = p->coop_grid_limit;
            uint32_t sg = ++c->seq[CH_GATHER], sm = ++c->seq[CH_MASK];
            CommDev cdv = cd;
            unsigned long long* rt = p->d_rank_ties; uint32_t* ties = p->d_chunk_ties; int64_t cb = chunk_begin, ce = chunk_end;
            void* args[] = {(void*)&a, (void*)&cdv, (void*)&sg, (void*)&sm, (void*)&rt, (void*)&ties, (void*)&em, (void*)&cb, (void*)&ce};
            B200P_CUDA(cudaLaunchCooperativeKernel((const void*)k_sharded_tail, dim3(grid), dim3(kThreads), args, 0, st));
            p->prov_armed = false; p->prov_target = nullptr;
            return B200P_OK;
        }
    }
    // B
    if (stages & B200P_SHARD_FINISH) {
        const int64_t cblocks = (p->cand_capacity + 8 * kThreads - 1) / (8 * kThreads);
        k_sharded_finish<<<p->grid_for(cblocks, 1), kThreads, 0, st>>>(a, cd, ++c->seq[CH_GATHER], p->d_rank_ties);
        B200P_LAUNCH_CHECK("k_sharded_finish");
    }
    // T
    if ((stages & B200P_SHARD_TIES) && mode == B200P_MODE_EXACT_K) {
        rc = launch_tie_count(p, key_source, d_old_mask, chunk_begin, chunk_end, st); if (rc) return rc;
        k_tie_scan<<<1, 1024, 0, st>>>(p->d_state, p->d_chunk_ties, chunk_begin, chunk_end, 0ull, p->d_rank_ties, c->rank);
        B200P_LAUNCH_CHECK("k_tie_scan");
    }
    // E, P
    if (stages & B200P_SHARD_EMIT) {
        rc = b200p_emit_masks(p, key_source, mode, 0, 0.f, d_old_mask, d_mask, 0, chunk_begin, chunk_end, stream); if (rc) return rc;
    }
    if (stages & B200P_SHARD_PUSH) return b200p_comm_mask_allgather(c, p, chunk_begin, chunk_end, stream);
    return B200P_OK;
}

// ---- fused SNIP mask build ------------------------------------------------------------------------
template <bool ACC>
static cudaError_t launch_snip_sample(int nb, int grid, cudaStream_t st, const SampleArgs& sa, ChunkTab w, const GradTabs& g, ChunkTab s) {
    switch (nb) {
        case 1: return launch_kernel(k_snip_sample<ACC, 1>, grid, st, sa, w, g, s);
        case 2: return launch_kernel(k_snip_sample<ACC, 2>, grid, st, sa, w, g, s);
        case 3: return launch_kernel(k_snip_sample<ACC, 3>, grid, st, sa, w, g, s);
        case 4: return launch_kernel(k_snip_sample<ACC, 4>, grid, st, sa, w, g, s);
        case 5: return launch_kernel(k_snip_sample<ACC, 5>, grid, st, sa, w, g, s);
        case 6: return launch_kernel(k_snip_sample<ACC, 6>, grid, st, sa, w, g, s);
        case 7: return launch_kernel(k_snip_sample<ACC, 7>, grid, st, sa, w, g, s);
        default: return launch_kernel(k_snip_sample<ACC, 8>, grid, st, sa, w, g, s);
    }
}
template <bool ACC>
static cudaError_t launch_snip_refresh_sample(int nb, int grid, cudaStream_t st, const SampleArgs& sa, ChunkTab w, const GradTabs& g, ChunkTab s, const RefreshArgs& rf) {
    switch (nb) {
        case 1: return launch_kernel(k_snip_refresh_sample<ACC, 1>, grid, st, sa, w, g, s, rf);
        case 2: return launch_kernel(k_snip_refresh_sample<ACC, 2>, grid, st, sa, w, g, s, rf);
        case 3: return launch_kernel(k_snip_refresh_sample<ACC, 3>, grid, st, sa, w, g, s, rf);
        case 4: return launch_kernel(k_snip_refresh_sample<ACC, 4>, grid, st, sa, w, g, s, rf);
        case 5: return launch_kernel(k_snip_refresh_sample<ACC, 5>, grid, st, sa, w, g, s, rf);
        case 6: return launch_kernel(k_snip_refresh_sample<ACC, 6>, grid, st, sa, w, g, s, rf);
        case 7: return launch_kernel(k_snip_refresh_sample<ACC, 7>, grid, st, sa, w, g, s, rf);
        default: return launch_kernel(k_snip_refresh_sample<ACC, 8>, grid, st, sa, w, g, s, rf);
    }
}
template <bool ACC>
static cudaError_t launch_snip_sweep(int nb, int grid, cudaStream_t st, const PassArgs& a, ChunkTab w, const GradTabs& g, int vec) {
    switch (nb) {
        case 1: return launch_kernel(k_snip_score_sweep<ACC, 1>, grid, st, a, w, g, vec);
        case 2: return launch_kernel(k_snip_score_sweep<ACC, 2>, grid, st, a, w, g, vec);
        case 3: return launch_kernel(k_snip_score_sweep<ACC, 3>, grid, st, a, w, g, vec);
        case 4: return launch_kernel(k_snip_score_sweep<ACC, 4>, grid, st, a, w, g, vec);
        case 5: return launch_kernel(k_snip_score_sweep<ACC, 5>, grid, st, a, w, g, vec);
        case 6: return launch_kernel(k_snip_score_sweep<ACC, 6>, grid, st, a, w, g, vec);
        case 7: return launch_kernel(k_snip_score_sweep<ACC, 7>, grid, st, a, w, g, vec);
        default: return launch_kernel(k_snip_score_sweep<ACC, 8>, grid, st, a, w, g, vec);
    }
}

// refresh: segment pointers of the (new) gradient tensors behind g_tables; the sample kernel writes the tables itself
static int snip_score_select_impl(b200p_plan* p, const b200p_ptrtable* const* g_tables, int n_sets, uint64_t k,
                                  uint32_t* d_prov_target, void* stream, const RefreshArgs* refresh);
extern "C" int b200p_snip_score_select(b200p_plan* p, const b200p_ptrtable* const* g_tables, int n_sets, uint64_t k,
                                       uint32_t* d_prov_target, void* stream) {
    return snip_score_select_impl(p, g_tables, n_sets, k, d_prov_target, stream, nullptr);
}
static int snip_score_select_impl(b200p_plan* p, const b200p_ptrtable* const* g_tables, int n_sets, uint64_t k,
                                  uint32_t* d_prov_target, void* stream, const RefreshArgs* refresh) {
    B200P_REQUIRE(p != nullptr && g_tables != nullptr, B200P_EINVAL, "snip_mask_build: null argument");
    B200P_REQUIRE(n_sets >= 1, B200P_EINVAL, "snip_mask_build: need at least one gradient set");
    B200P_REQUIRE(p->bound[B200P_SLOT_W] && p->bound[B200P_SLOT_SCORE], B200P_ESTATE, "snip_mask_build: W and SCORE slots must be bound");
    B200P_REQUIRE(k >= 1 && k <= (uint64_t)p->total, B200P_EINVAL, "snip_mask_build: k must be in [1, N]");
    for (int i = 0; i < n_sets; ++i)
        B200P_REQUIRE(g_tables[i] != nullptr && g_tables[i]->plan == p, B200P_EINVAL, "snip_mask_build: table belongs to another plan");
    B200P_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    { int rc = coop_ctas(p); if (rc) return rc; }
    // all but the last group of up to 8 sets are plain accumulate launches; the last group is fused with the select
    const int last0 = ((n_sets - 1) / kMaxSets) * kMaxSets, nb = n_sets - last0;
    const bool unfused = p->select_impl == B200P_SELECT_EXACT || p->coop_ctas_per_sm < 0;
    const int n_plain = unfused ? n_sets : last0;
    if (n_plain > 0) { int rc = b200p_score_accumulate_multi(p, g_tables, n_plain, 0, 0, -1, stream); if (rc) return rc; }
    if (unfused) {
        p->prov_target = d_prov_target;
        return b200p_select_kth(p, B200P_KEY_SCORE, nullptr, k, B200P_MODE_SNIP_STRICT, stream);
    }
    const bool acc = last0 > 0;
    bool vec = p->vec_ok[B200P_SLOT_W] && p->vec_ok[B200P_SLOT_SCORE];
    GradTabs g;
    for (int b = 0; b < kMaxSets; ++b) { g.t[b] = (ChunkTab)g_tables[last0 + (b < nb ? b : 0)]->d_tab; if (b < nb) vec = vec && g_tables[last0 + b]->vec_ok; }
    // an emit that follows with the same arguments patches the provisional mask the sweep writes (into d_prov_target if given)
    p->prov_target = d_prov_target;
    p->prov_armed = true; p->prov_key_source = B200P_KEY_SCORE; p->prov_mode = B200P_MODE_SNIP_STRICT; p->prov_old_mask = nullptr;
    p->prov_c0 = 0; p->prov_c1 = p->n_chunks; p->emit_done = false;
    // S': sample
    SampleArgs sa;
    sa.chunk_n = p->d_chunk_n; sa.key_tab = p->tab(B200P_SLOT_SCORE); sa.old_mask = nullptr; sa.hist = p->d_hist; sa.st = p->d_state;
    sa.ticket = ticket_ptr(p); sa.n_chunks = p->n_chunks; sa.c_begin = 0; sa.comm_seq = 0u; sa.cache = nullptr; sa.vec_ok = vec ? 1 : 0;
    sa.k = k; sa.n_total = (unsigned long long)p->total; sa.mode = (uint32_t)B200P_MODE_SNIP_STRICT; sa.sigmas = 12u; sa.slot_shift = 4;
    const int64_t sblocks = (p->n_chunks * kSnipGranulesPerChunk + kThreads - 1) / kThreads;      // one granule per thread
    if (refresh) {
        if (acc) B200P_CUDA(launch_snip_refresh_sample<true>(nb, p->grid_for(sblocks, 4), st, sa, p->tab(B200P_SLOT_W), g, p->tab(B200P_SLOT_SCORE), *refresh));
        else     B200P_CUDA(launch_snip_refresh_sample<false>(nb, p->grid_for(sblocks, 4), st, sa, p->tab(B200P_SLOT_W), g, p->tab(B200P_SLOT_SCORE), *refresh));
    } else {
        if (acc) B200P_CUDA(launch_snip_sample<true>(nb, p->grid_for(sblocks, 4), st, sa, p->tab(B200P_SLOT_W), g, p->tab(B200P_SLOT_SCORE)));
        else     B200P_CUDA(launch_snip_sample<false>(nb, p->grid_for(sblocks, 4), st, sa, p->tab(B200P_SLOT_W), g, p->tab(B200P_SLOT_SCORE)));
    }
    B200P_LAUNCH_CHECK("k_snip_sample");
    // A': score + sweep
    PassArgs a;
    fill_pass_args(p, a, B200P_KEY_SCORE, nullptr, 0, p->n_chunks, 1, 0, k, B200P_MODE_SNIP_STRICT, 1, true);
    { int rc = plan_time_mark(p, 0, st); if (rc) return rc; }
    if (acc) B200P_CUDA(launch_snip_sweep<true>(nb, p->grid_for(p->n_chunks, 4), st, a, p->tab(B200P_SLOT_W), g, vec ? 1 : 0));
    else     B200P_CUDA(launch_snip_sweep<false>(nb, p->grid_for(p->n_chunks, 4), st, a, p->tab(B200P_SLOT_W), g, vec ? 1 : 0));
    B200P_LAUNCH_CHECK("k_snip_score_sweep");
    { int rc = plan_time_mark(p, 1, st); if (rc) return rc; }
    return launch_finish(p, a, B200P_KEY_SCORE, B200P_MODE_SNIP_STRICT, nullptr, st);
}

extern "C" int b200p_snip_mask_build(b200p_plan* p, const b200p_ptrtable* const* g_tables, int n_sets, uint64_t k,
                                     uint32_t* d_new_mask, void* stream) {
    B200P_REQUIRE(d_new_mask != nullptr && p != nullptr, B200P_EINVAL, "snip_mask_build: null argument");
    p->fuse_emit = true; p->emit_done = false;
    int rc = b200p_snip_score_select(p, g_tables, n_sets, k, d_new_mask, stream);
    p->fuse_emit = false;
    if (rc) { p->prov_target = nullptr; return rc; }
    if (p->emit_done) {                       // the finish kernel patched the mask itself
        p->emit_done = false; p->prov_armed = false; p->prov_target = nullptr;
        return B200P_OK;
    }
    return b200p_emit_masks(p, B200P_KEY_SCORE, B200P_MODE_SNIP_STRICT, 0, 0.f, nullptr, d_new_mask, 0, 0, -1, stream);
}

// b200p_snip_mask_build for gradient tensors that are new since the tables were filled (every real build: each backward pass
// allocates its gradients): h_ptrs[i][t] is the device pointer of segment t of gradient set i.  One launch less than
// b200p_ptrtables_update + b200p_snip_mask_build: the sample kernel writes the per-chunk tables on its way.  Falls back to
// that pair when the pointers do not fit the kernel arguments (n_sets * n_seg > 1024), when there are more than 8 sets, or
// when the fused path is not available (exact select, no cooperative launch).
extern "C" int b200p_snip_mask_build_refresh(b200p_plan* p, b200p_ptrtable* const* g_tables, const void* const* const* h_ptrs, int n_sets,
                                             uint64_t k, uint32_t* d_new_mask, void* stream) {
    B200P_REQUIRE(p != nullptr && g_tables != nullptr && h_ptrs != nullptr && d_new_mask != nullptr, B200P_EINVAL, "snip_mask_build_refresh: null argument");
    B200P_REQUIRE(n_sets >= 1, B200P_EINVAL, "snip_mask_build_refresh: need at least one gradient set");
    for (int i = 0; i < n_sets; ++i)
        B200P_REQUIRE(g_tables[i] != nullptr && g_tables[i]->plan == p && h_ptrs[i] != nullptr, B200P_EINVAL, "snip_mask_build_refresh: tables must belong to this plan");
    { int rc = coop_ctas(p); if (rc) return rc; }
    const bool fused_ok = n_sets <= kMaxSets && (long long)n_sets * p->n_seg <= kMultiPtrs && p->select_impl != B200P_SELECT_EXACT &&
                          p->coop_ctas_per_sm > 0 && k >= 1 && k <= (uint64_t)p->total;
    if (!fused_ok) {
        int rc = b200p_ptrtables_update(g_tables, h_ptrs, n_sets, B200P_SLOT_G, stream); if (rc) return rc;
        return b200p_snip_mask_build(p, (const b200p_ptrtable* const*)g_tables, n_sets, k, d_new_mask, stream);
    }
    RefreshArgs rf;
    rf.chunk_seg = p->d_chunk_seg; rf.chunk_elem0 = p->d_chunk_elem0; rf.n_seg = p->n_seg;
    for (int i = 0; i < n_sets; ++i) {
        bool vec = true;
        for (int t = 0; t < p->n_seg; ++t) {
            const void* q = h_ptrs[i][t];
            B200P_REQUIRE(q != nullptr, B200P_EINVAL, "snip_mask_build_refresh: null segment pointer");
            rf.pack.p[i * p->n_seg + t] = const_cast<void*>(q);
            if ((uintptr_t)q & 15u) vec = false;
        }
        g_tables[i]->vec_ok = vec;
        for (int sl = 0; sl < B200P_NUM_SLOTS; ++sl)
            if (p->d_tab[sl] == g_tables[i]->d_tab) p->vec_ok[sl] = vec;
    }
    p->fuse_emit = true; p->emit_done = false;
    int rc = snip_score_select_impl(p, (const b200p_ptrtable* const*)g_tables, n_sets, k, d_new_mask, stream, &rf);
    p->fuse_emit = false;
    if (rc) { p->prov_target = nullptr; return rc; }
    if (p->emit_done) {
        p->emit_done = false; p->prov_armed = false; p->prov_target = nullptr;
        return B200P_OK;
    }
    return b200p_emit_masks(p, B200P_KEY_SCORE, B200P_MODE_SNIP_STRICT, 0, 0.f, nullptr, d_new_mask, 0, 0, -1, stream);
}

extern "C" int b200p_select_last_trace(uint64_t* h_out16) {
    B200P_REQUIRE(h_out16 != nullptr, B200P_EINVAL, "select_last_trace: null output");
    B200P_CUDA(cudaDeviceSynchronize());
    B200P_CUDA(cudaMemcpyFromSymbol(h_out16, g_sel_stamps, sizeof(unsigned long long) * 16));
    unsigned long long zero[16] = {0};
    B200P_CUDA(cudaMemcpyToSymbol(g_sel_stamps, zero, sizeof(zero)));
    return B200P_OK;
}

extern "C" int b200p_select_result(b200p_plan* p, b200p_select_result_t* h_out, void* stream) {
    B200P_REQUIRE(p != nullptr && h_out != nullptr, B200P_EINVAL, "select_result: null argument");
    B200P_CUDA(cudaSetDevice(p->device));
    SelState s;
    B200P_CUDA(cudaMemcpyAsync(&s, p->d_state, sizeof(s), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    B200P_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    h_out->k = s.k_request; h_out->n_valid = s.n_valid; h_out->n_less = s.n_less; h_out->n_equal = s.n_equal;
    h_out->quota = s.quota; h_out->n_kept = s.n_kept; h_out->threshold = s.threshold; h_out->thr_key = s.thr_key;
    h_out->passes_full = s.passes_full ? s.passes_full : (s.collect ? 2u : 3u); h_out->collected = s.collect ? s.cand_count : 0u;
    h_out->miss = s.miss; h_out->reserved_ = 0u;
    return B200P_OK;
}
