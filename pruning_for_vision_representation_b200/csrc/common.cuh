// common.cuh — shared device helpers and the plan object of the B200 pruning hot path.
// sm_100a only.  See include/b200prune.h for the data model (segments, chunks, packed masks).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/b200prune.h"

namespace b200p {

constexpr int kChunk   = B200P_CHUNK;            // elements per chunk
constexpr int kThreads = 256;                    // threads per CTA of the streaming kernels
constexpr int kVecPerThread = kChunk / (4 * kThreads);   // float4 per thread per chunk = 4
constexpr int kWordsPerChunk = B200P_WORDS_PER_CHUNK;
constexpr int kHistBins = 4096;                  // 12-bit digits
constexpr int kHistExtra = 8;                    // scalar counters behind the bins: [0] alive keys (sample pass), [1] keys below the bracket,
                                                 // [2] NaN keys cleared from the provisional mask, [7] ticket of the mask push
constexpr int kHistStride = kHistBins + kHistExtra;   // u64 words of one copy of the global histogram
constexpr int kHistReplicas = 4;                 // copies of the global histogram the sample / sweep kernels spread their flush over (select.cu: flush_hist);
                                                 // the scalar counters and every other user live in copy 0
constexpr int kTieListCap = 512;                 // short list of tied candidates behind the per-chunk tie table (select.cu: tie_list_gather)
constexpr uint32_t kNanKey = 0x7FFFFFFFu;        // every NaN sorts last (torch.sort semantics)

// digit layout of the 31-bit key: pass 0 -> bits 30..19, pass 1 -> bits 18..7, pass 2 -> bits 6..0
__host__ __device__ __forceinline__ uint32_t digit_of(uint32_t key, int pass) {
    return pass == 0 ? (key >> 19) : pass == 1 ? ((key >> 7) & 0xFFFu) : (key & 0x7Fu);
}
__host__ __device__ __forceinline__ int digit_bins(int pass) { return pass == 2 ? 128 : 4096; }
// mask selecting the bits already fixed before `pass`
__host__ __device__ __forceinline__ uint32_t prefix_mask_before(int pass) {
    return pass == 0 ? 0u : pass == 1 ? 0x7FF80000u : 0x7FFFFF80u;
}

// key of a score / weight: bit pattern of |x|, NaN canonicalised to the maximum key.
// For non-negative fp32 the integer order of the bit patterns is the float order.
__device__ __forceinline__ uint32_t key_of(float x) {
    uint32_t u = __float_as_uint(x) & 0x7FFFFFFFu;
    return u > 0x7F800000u ? kNanKey : u;
}
__host__ __device__ __forceinline__ float key_to_float(uint32_t key) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(key == kNanKey ? 0x7FC00000u : key);
#else
    union { uint32_t u; float f; } c; c.u = (key == kNanKey ? 0x7FC00000u : key); return c.f;
#endif
}

// Per-chunk tables.  A kernel never walks segment tables: for chunk c it reads the number of valid
// elements chunk_n[c] and, per slot it touches, the address of the chunk's first element
// tab[c] — independent loads that are issued one iteration ahead (ChunkCursor), so the
// grid-stride loop has no dependent pointer chase in front of its data loads.
typedef void* const* ChunkTab;          // [n_chunks] device pointers, one table per slot

template <typename T>
__device__ __forceinline__ T* chunk_ptr(ChunkTab tab, int64_t c) {
    return reinterpret_cast<T*>(__ldg(reinterpret_cast<const unsigned long long*>(tab) + c));
}

// segment pointers that travel as kernel arguments (b200p_ptrtables_update; the SNIP sample kernel that refreshes the gradient
// tables itself): table t's pointer of segment g sits at p[t * n_seg + g]
constexpr int kMultiPtrs = 1024, kMultiTabs = 16;
struct PtrPackBig { void* p[kMultiPtrs]; };

// up to 8 gradient sets (one per-chunk pointer table each) folded by one score launch
constexpr int kMaxSets = 8;
struct GradTabs { ChunkTab t[kMaxSets]; };

// 128-bit streaming loads / stores.  Read-only streams go through the non-coherent path
// without allocating in L1; read-modify-write streams use plain (coherent) accesses.
__device__ __forceinline__ float4 ld_nc_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_f4(float* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Combine the 4-bit nibbles of 8 consecutive lanes into one 32-bit mask word.
// Lane L contributes bits [4*(L&7), 4*(L&7)+4).  Every lane of the group of 8 returns the word.
__device__ __forceinline__ uint32_t gather_nibbles(uint32_t nib) {
    uint32_t v = (nib & 0xFu) << (4 * (threadIdx.x & 7));
    v |= __shfl_xor_sync(0xFFFFFFFFu, v, 1);
    v |= __shfl_xor_sync(0xFFFFFFFFu, v, 2);
    v |= __shfl_xor_sync(0xFFFFFFFFu, v, 4);
    return v;
}
// nibble of this thread out of a mask word shared by its group of 8 lanes
__device__ __forceinline__ uint32_t nibble_of(uint32_t word) {
    return (word >> (4 * (threadIdx.x & 7))) & 0xFu;
}
// In the vector path thread `tid` owns float4 number v = j*kThreads + tid of the chunk
// (elements 4v..4v+3); the mask word holding them is word v/8 = j*32 + tid/8.
__device__ __forceinline__ int vec_word_index(int j) { return j * (kThreads / 8) + (threadIdx.x >> 3); }

// select state kept in the plan workspace (device).  Host mirror: b200p_select_result_t.
struct SelState {
    unsigned long long k;            // 1-based rank still to find inside the current bucket
    unsigned long long k_request;    // rank as requested
    unsigned long long n_valid;      // alive keys
    unsigned long long n_less;       // alive keys below the current bucket
    unsigned long long n_equal;      // final: keys equal to the threshold
    unsigned long long quota;        // final: tied keys to prune (EXACT_K)
    unsigned long long n_kept;       // kept bits of the last emit
    unsigned long long bucket_count; // population of the bucket chosen by the last scan
    unsigned long long tie_seen;     // ties located in chunks before tie_chunk (tie pass)
    long long          tie_chunk;    // chunk in which the quota runs out (-1: every tie pruned)
    uint32_t prefix;                 // key bits fixed so far
    uint32_t thr_key;                // final key
    float    threshold;              // final fp32 threshold
    uint32_t mode;                   // B200P_MODE_*
    uint32_t collect;                // 1: pass-1 bucket was gathered into the candidate buffer
    uint32_t allow_collect;
    uint32_t cand_count;             // candidates gathered
    uint32_t tie_resid;              // ties to prune inside tie_chunk
    uint32_t need_ties;              // EXACT_K and quota < n_equal
    uint32_t passes_full;            // full-data passes executed
    // sampled (bracketed) select
    uint32_t sample_ok;              // the sample pass produced a usable bracket
    uint32_t miss;                   // the k-th key was not inside the bracket (or it overflowed): exact fallback ran
    uint32_t lo_bucket, hi_bucket;   // bracket in units of the 12-bit digit (inclusive)
    uint32_t win_lo;                 // first key of the 1024-key window chosen by the bracket pass
    uint32_t prov_ok;                // the provisional mask (bits: alive && key >= bracket base) and the candidate list are valid
    uint32_t pad_[2];                // [0] ticket of the last-CTA pattern, [1] alive NaN keys pruned by the provisional mask (SNIP_STRICT)
};

}  // namespace b200p

// ---------------------------------------------------------------------------------------
// plan object (host)
struct b200p_plan {
    int device = 0;
    int n_seg = 0;
    int64_t total = 0;
    int64_t n_chunks = 0;
    int64_t cand_capacity = 0;
    int num_sms = 148;
    int coop_ctas_per_sm = 0;      // co-resident CTAs of the cooperative finish kernel (0: not queried yet)
    int coop_grid_limit = 0;       // B200P_OPT_COOP_GRID: cap on the grid of cooperative launches (several plans sharing one device)
    int select_impl = 0;           // 0 sampled (bracketed) select, 1 exact 3-pass radix select
    std::vector<int64_t> numel;
    std::vector<int64_t> seg_chunk_start;   // n_seg + 1
    std::vector<int64_t> seg_flat_start;    // n_seg + 1
    // device tables
    int32_t* d_chunk_seg = nullptr;         // [n_chunks] segment of each chunk (bind kernel only)
    int32_t* d_chunk_n = nullptr;           // [n_chunks] valid elements of each chunk (1..kChunk)
    int64_t* d_chunk_elem0 = nullptr;       // [n_chunks] first element of the chunk inside its segment
    void**   d_tab_own[B200P_NUM_SLOTS] = {nullptr};   // plan-owned per-chunk pointer tables
    void**   d_tab[B200P_NUM_SLOTS] = {nullptr};       // table currently bound to each slot
    bool     bound[B200P_NUM_SLOTS] = {false};
    bool     vec_ok[B200P_NUM_SLOTS] = {false};
    // workspace
    unsigned long long* d_hist = nullptr;   // kHistBins
    b200p::SelState*    d_state = nullptr;
    uint32_t* d_cand_key = nullptr;         // cand_capacity
    uint32_t* d_cand_pos = nullptr;         // cand_capacity   (chunk * kChunk + element)
    uint32_t* d_chunk_ties = nullptr;       // n_chunks
    uint32_t* d_prov = nullptr;             // n_chunks * 128 words: provisional packed mask written by the select sweep
    // emit-by-patch: valid for the emit that directly follows b200p_select_kth with the same arguments
    uint32_t* prov_target = nullptr;        // where the sweep writes the provisional mask (nullptr: d_prov)
    bool prov_armed = false; int prov_key_source = -1; int prov_mode = -1; const uint32_t* prov_old_mask = nullptr;
    // B200P_OPT_REUSE_SAMPLE: cached sample histogram of the last sampled select
    bool reuse_sample = false, sample_cache_valid = false; int sample_cache_key = -1; const uint32_t* sample_cache_mask = nullptr;
    int64_t sample_cache_c0 = 0, sample_cache_c1 = 0; const void* sample_cache_tab = nullptr;
    unsigned long long* d_sample_cache = nullptr;
    bool fuse_emit = false;                 // b200p_mask_build / b200p_snip_mask_build: the finish kernel emits the mask itself
    bool emit_done = false;                 // ... and did (the caller skips its emit launch)
    int64_t prov_c0 = 0, prov_c1 = 0;       // chunk range the sweep covered (the whole set on one GPU, the rank's slice when sharded)
    unsigned long long* d_rank_ties = nullptr;   // [8] sharded select: every rank's count of the threshold key (k_sharded_finish)
    // lazily created arena for the host-buffer entry points
    float* arena_w = nullptr; float* arena_g[2] = {nullptr, nullptr}; float* arena_score = nullptr;
    uint32_t* arena_mask = nullptr; uint32_t* arena_old_mask = nullptr;
    struct b200p_ptrtable* arena_gtab[2] = {nullptr, nullptr};
    cudaStream_t arena_streams[2] = {nullptr, nullptr};
    cudaEvent_t  arena_events[4] = {nullptr, nullptr, nullptr, nullptr};

    // B200P_OPT_TIME_SWEEP: CUDA events around every k_snip_score_sweep launch (ring of pairs), read back with
    // b200p_plan_kernel_time_ms — how bench.py times the dominant kernel of the fused sequence on its own stream
    bool time_sweep = false;
    std::vector<cudaEvent_t> tev;           // 2 * kTimedPairs events, created on first use
    int tev_head = 0, tev_pending = 0;      // next pair to record / recorded pairs not yet folded into the sum
    double tev_sum_ms = 0.0; long long tev_n = 0;

    template <typename T = void> b200p::ChunkTab tab(int slot) const { return (b200p::ChunkTab)d_tab[slot]; }
    int grid_for(int64_t chunks, int ctas_per_sm) const {
        int64_t g = (int64_t)num_sms * ctas_per_sm;
        if (g > chunks) g = chunks;
        return g < 1 ? 1 : (int)g;
    }
};

// reusable per-chunk pointer table (b200p_ptrtable_create); binding it is a host-side swap
struct b200p_ptrtable {
    b200p_plan* plan = nullptr;
    void** d_tab = nullptr;
    bool vec_ok = false;
};

namespace b200p {
void set_error(const std::string& msg);
int  plan_time_mark(b200p_plan* p, int which, cudaStream_t st);
int  cuda_fail(cudaError_t e, const char* what);
}  // namespace b200p
struct b200p_comm;
extern "C" int b200p_comm_mask_allgather(b200p_comm* c, b200p_plan* p, int64_t chunk_begin, int64_t chunk_end, void* stream);

#define B200P_CUDA(call)                                                        \
    do { cudaError_t e__ = (call);                                              \
         if (e__ != cudaSuccess) return b200p::cuda_fail(e__, #call); } while (0)
#define B200P_REQUIRE(cond, code, msg)                                          \
    do { if (!(cond)) { b200p::set_error(msg); return (code); } } while (0)
#define B200P_LAUNCH_CHECK(name)                                                \
    do { cudaError_t e__ = cudaGetLastError();                                  \
         if (e__ != cudaSuccess) return b200p::cuda_fail(e__, name); } while (0)
