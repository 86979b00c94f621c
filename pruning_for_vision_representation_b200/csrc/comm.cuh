// comm.cuh — peer-memory window of one rank and the in-kernel collectives built on it.
//
// the sharded mask build needs three tiny exchanges (sample histogram, bracket histogram + below-count,
// window histogram) and one bulk one (packed mask words; partial SNIP scores).  Over NVLink 5 / NVSwitch every GPU can
// store into every peer's memory, so none of them is an NCCL call here: each rank exposes one WINDOW (cudaMalloc'd,
// opened by the peers through CUDA IPC, or plain pointers when the "ranks" are plans on one device), and the kernel that
// produces the data pushes it straight into the peers' windows and raises a flag there.  The last CTA of the select's
// sample / sweep kernels does the all-reduce itself, between its own histogram flush and its own scan: no launch, no
// host round trip, no separate reduction kernel.
//
// Protocol (every rank runs the same sequence of operations, SPMD):
//   * an operation on channel ch carries a sequence number seq = 1, 2, 3, ... (host-side counter per channel, passed as
//     a kernel argument); payload slots are double-buffered by seq & 1;
//   * writer r: stores its payload into slot [seq & 1][r] of EVERY rank's window (its own included), then
//     __threadfence_system(), then st.release.sys of seq into flag[ch][r] of every window;
//   * reader: spins (bounded) on ld.acquire.sys of its OWN window's flag[ch][p] >= seq for all p, then reads the slots.
//   A rank can be at most one operation ahead of the slowest one on a channel (it needs everybody's seq before it can
//   finish seq), so slot seq & 1 is never overwritten while somebody still reads seq - 2's data... which every rank
//   finished before it wrote seq - 1.
//   * a spin that runs out (~2 s) sets CommDev::err_flag and returns false: a broken peer shows up as an error code,
//     never as a hung GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace b200p {

constexpr int kCommMaxWorld = 8;
// channels
enum { CH_HIST = 0, CH_GATHER = 1, CH_MASK = 2, CH_BARRIER = 3, CH_COUNT = 4 };
constexpr int kCommHistBins = 4096, kCommHistExtra = 8;            // u32 bins + u64 extras per rank slot
constexpr int kCommGatherWords = 1024 + 8;                         // u32 window histogram + scalars per rank slot

// byte offsets inside a window (identical on every rank)
struct CommLayout {
    long long flags;        // u32 [CH_COUNT][kCommMaxWorld]
    long long err;          // u32 error flag (own window only)
    long long trace;        // u64 [CH_COUNT][2][8] globaltimer stamps of the last collective per channel and seq parity (own window only)
    long long hist_bins;    // u32 [2][world][kCommHistBins]
    long long hist_extra;   // u64 [2][world][kCommHistExtra]
    long long gather;       // u32 [2][world][kCommGatherWords]
    long long mask;         // u32 [mask_words]: the full packed mask, every rank's slice pushed in by its owner
    long long score;        // f32 [world][score_cap]: partial SNIP scores of THIS rank's slice, one part per source rank
    long long total;
};

struct CommDev {
    char* win[kCommMaxWorld];      // mapped base of every rank's window (win[rank] is the own one)
    CommLayout lay;
    long long score_cap;           // elements per part of the score area
    int rank, world;
    int push_lsu;                  // 1 (default): mask push with LSU stores; 0: through the bulk-copy engine (B200P_PUSH_BULK=1)
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t* comm_flag(const CommDev& c, int dst_rank, int ch, int src_rank) {
    return reinterpret_cast<uint32_t*>(c.win[dst_rank] + c.lay.flags) + ch * kCommMaxWorld + src_rank;
}

// Called by ALL threads of one CTA after they have stored their part of the payload into the peers' windows:
// publishes seq to every rank, then waits for every rank's seq.  Returns false on a time-out (err flag set).
__device__ __forceinline__ unsigned long long comm_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// measurement aid (b200p_comm_trace): [0] payload stores issued [1] own fence done [2] own flag stored [3] last peer's flag seen
// [4] leaving; written by the signalling threads of the last collective on (channel, seq parity)
__device__ __forceinline__ unsigned long long* comm_trace_slot(const CommDev& c, int ch, uint32_t seq) {
    return reinterpret_cast<unsigned long long*>(c.win[c.rank] + c.lay.trace) + (ch * 2 + (int)(seq & 1u)) * 8;
}
__device__ __forceinline__ bool comm_signal_and_wait(const CommDev& c, int ch, uint32_t seq) {
    __shared__ int s_comm_ok;
    unsigned long long* tr = comm_trace_slot(c, ch, seq);
    if (threadIdx.x == 0) { s_comm_ok = 1; tr[3] = 0ull; }
    __syncthreads();                       // every payload store of the CTA is issued (CTA-scope happens-before)
    if (threadIdx.x == 0) tr[0] = comm_globaltimer();
    if ((int)threadIdx.x < c.world) {
        // ONE system-scope fence per signalling thread, after the barrier: fences are cumulative, so it also orders the
        // stores of the other threads it synchronised with (256 concurrent fence.sys cost ~15 us per collective)
        __threadfence_system();
        if (threadIdx.x == 0) tr[1] = comm_globaltimer();
        st_release_sys_u32(comm_flag(c, threadIdx.x, ch, c.rank), seq);
        if (threadIdx.x == 0) tr[2] = comm_globaltimer();
        const uint32_t* f = comm_flag(c, c.rank, ch, threadIdx.x);
        bool ok = false;
        for (unsigned spin = 0; spin < (1u << 24); ++spin) {
            // sequence numbers only grow; compare as a signed distance so that a wrap is harmless
            if ((int32_t)(ld_acquire_sys_u32(f) - seq) >= 0) { ok = true; break; }
            if (spin > 64) __nanosleep(128);
        }
        atomicMax(tr + 3, comm_globaltimer());
        if (!ok) { s_comm_ok = 0; *reinterpret_cast<volatile uint32_t*>(c.win[c.rank] + c.lay.err) = 1u + (uint32_t)ch; }
    }
    __syncthreads();
    if (threadIdx.x == 0) tr[4] = comm_globaltimer();
    return s_comm_ok != 0;
}

// Copy the packed-mask words [w_begin, w_end) of the own window into every peer's window (grid-stride over the launching
// grid; chunk ranges start on multiples of 128 words); each CTA ends with one cumulative system fence.
// The all-gather is bound by NVLink ingress (every rank receives (G-1)/G of the mask: 25 MB for ViT-L/16 at 8 GPUs, 28 us at
// 900 GB/s).  LSU stores — 16-byte pieces, four in flight per lane, every piece to every peer — reach 485-550 GB/s (51 us at
// 8 GPUs, 26 us for 14 MB to one peer at 2 GPUs; b200p_comm_trace).  Tried: tiles staged in shared memory and handed to the
// bulk-copy engine (cp.async.bulk global <- shared, one instruction per 8 KB tile and peer, two tiles in flight) — bit-identical
// and exactly as fast, so the rate is the link's, not the store path's.  B200P_PUSH_BULK=1 (read at comm creation) selects it.
constexpr int kPushTileQ = 512;                                     // 16-byte pieces per tile (8 KB)
__device__ __forceinline__ void comm_push_mask_words(const CommDev& c, long long w_begin, long long w_end) {
    const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(c.win[c.rank] + c.lay.mask);
    const long long q0 = w_begin >> 2, q1 = w_end >> 2;
    if (c.world > 1 && c.push_lsu) {
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long q = q0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += 4 * stride) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) if (q + u * stride < q1) v[u] = __ldcg(reinterpret_cast<const uint4*>(src) + q + u * stride);
            for (int p = 0; p < c.world; ++p) {
                if (p == c.rank) continue;
                uint4* dst = reinterpret_cast<uint4*>(c.win[p] + c.lay.mask);
#pragma unroll
                for (int u = 0; u < 4; ++u) if (q + u * stride < q1) dst[q + u * stride] = v[u];
            }
        }
    } else if (c.world > 1) {
        __shared__ __align__(128) uint4 s_push[2][kPushTileQ];
        const long long n_tiles = (q1 - q0 + kPushTileQ - 1) / kPushTileQ;
        int buf = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, buf ^= 1) {
            const long long qa = q0 + t * kPushTileQ;
            const int cnt = (int)((q1 - qa) < kPushTileQ ? (q1 - qa) : kPushTileQ);
            // the copies issued from this buffer two tiles ago have read it (at most one newer group may still be reading the other)
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
            for (int i = threadIdx.x; i < cnt; i += blockDim.x) s_push[buf][i] = __ldcg(reinterpret_cast<const uint4*>(src) + qa + i);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the bulk-copy engine
            __syncthreads();
            if (threadIdx.x == 0) {
                const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(&s_push[buf][0]);
                for (int p = 0; p < c.world; ++p) {
                    if (p == c.rank) continue;
                    const uint4* dst = reinterpret_cast<const uint4*>(c.win[p] + c.lay.mask) + qa;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(saddr), "r"(cnt * 16) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // every bulk write of this CTA has been performed
    }
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
    __syncthreads();
}

// All-reduce (sum) of a 4096-bin u64 histogram + 8 u64 extras that lives in this rank's global memory, by ONE CTA of
// kThreads threads (the last CTA of the kernel that produced it).  Per-rank counts fit 32 bits (< 2^32 keys per rank).
// On return `hist` holds the sums on every rank (bit-identical: integer adds).
__device__ __forceinline__ bool comm_allreduce_hist(const CommDev& c, uint32_t seq, unsigned long long* __restrict__ hist) {
    const int tid = threadIdx.x, nt = blockDim.x, slot = seq & 1u;
    if (tid == 0) comm_trace_slot(c, CH_HIST, seq)[5] = comm_globaltimer();          // [5] all-reduce entered, [6] sums written
    // pack own bins to u32 once, then push to every window
    for (int b4 = tid; b4 < kCommHistBins / 4; b4 += nt) {
        uint4 v;
        v.x = (uint32_t)((volatile unsigned long long*)hist)[4 * b4 + 0]; v.y = (uint32_t)((volatile unsigned long long*)hist)[4 * b4 + 1];
        v.z = (uint32_t)((volatile unsigned long long*)hist)[4 * b4 + 2]; v.w = (uint32_t)((volatile unsigned long long*)hist)[4 * b4 + 3];
        for (int p = 0; p < c.world; ++p) {
            uint4* dst = reinterpret_cast<uint4*>(c.win[p] + c.lay.hist_bins) + ((size_t)(slot * c.world + c.rank) * (kCommHistBins / 4) + b4);
            *dst = v;
        }
    }
    if (tid < kCommHistExtra) {
        const unsigned long long e = ((volatile unsigned long long*)hist)[kCommHistBins + tid];
        for (int p = 0; p < c.world; ++p)
            reinterpret_cast<unsigned long long*>(c.win[p] + c.lay.hist_extra)[(size_t)(slot * c.world + c.rank) * kCommHistExtra + tid] = e;
    }
    if (!comm_signal_and_wait(c, CH_HIST, seq)) return false;
    // loads first, stores afterwards: a store into `hist` inside the loop orders every later load behind it and the 16
    // rounds of L2 latency run one after the other (b200p_comm_trace: 9-10 us for this sum, 2 GPUs)
    const uint4* bins4 = reinterpret_cast<const uint4*>(c.win[c.rank] + c.lay.hist_bins) + (size_t)slot * c.world * (kCommHistBins / 4);
    for (int b4 = tid; b4 < kCommHistBins / 4; b4 += nt) {          // nt = 256: four rounds of `world` 16-byte loads each
        uint4 v[kCommMaxWorld];
#pragma unroll
        for (int r = 0; r < kCommMaxWorld; ++r) v[r] = r < c.world ? __ldcg(bins4 + (size_t)r * (kCommHistBins / 4) + b4) : make_uint4(0u, 0u, 0u, 0u);
        unsigned long long s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
        for (int r = 0; r < kCommMaxWorld; ++r) { s0 += v[r].x; s1 += v[r].y; s2 += v[r].z; s3 += v[r].w; }
        hist[4 * b4 + 0] = s0; hist[4 * b4 + 1] = s1; hist[4 * b4 + 2] = s2; hist[4 * b4 + 3] = s3;
    }
    if (tid < kCommHistExtra) {
        const unsigned long long* ex = reinterpret_cast<const unsigned long long*>(c.win[c.rank] + c.lay.hist_extra) + (size_t)slot * c.world * kCommHistExtra;
        unsigned long long s = 0;
        for (int r = 0; r < c.world; ++r) s += __ldcg(ex + (size_t)r * kCommHistExtra + tid);
        hist[kCommHistBins + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) comm_trace_slot(c, CH_HIST, seq)[6] = comm_globaltimer();
    return true;
}

}  // namespace b200p

// host object
struct b200p_comm {
    int device = 0, rank = 0, world = 1;
    char* window = nullptr;                 // own window (cudaMalloc)
    b200p::CommLayout lay{};
    long long mask_words = 0, score_cap = 0;
    char* peers[b200p::kCommMaxWorld] = {nullptr};
    bool opened_ipc[b200p::kCommMaxWorld] = {false};
    bool connected = false;
    uint32_t seq[b200p::CH_COUNT] = {0, 0, 0, 0};      // last sequence number used per channel
    b200p::CommDev dev() const {
        b200p::CommDev d;
        for (int i = 0; i < b200p::kCommMaxWorld; ++i) d.win[i] = peers[i];
        d.lay = lay; d.score_cap = score_cap; d.rank = rank; d.world = world;
        static const int lsu = [] { const char* e = getenv("B200P_PUSH_BULK"); return e && atoi(e) == 1 ? 0 : 1; }();
        d.push_lsu = lsu;
        return d;
    }
};
