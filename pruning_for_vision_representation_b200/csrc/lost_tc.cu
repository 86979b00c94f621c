// lost_tc.cu — K6 on the 5th-generation tensor cores: batched Gram matrix A_b = F_b F_b^T
// (object_discovery.py:39) as a TMA -> tcgen05.mma (kind::tf32) -> TMEM pipeline with fp32-grade
// accuracy from the 3xTF32 split, and the degree count of patch_scoring (:77-87) fused into the
// TMEM epilogue.
//
//   k_lost_split_tf32   F -> (F_hi, F_lo):  hi = tf32(f) (round to nearest), lo = tf32(f - hi).
//                       hi + lo carries 21+ mantissa bits; A = hi.hi^T + hi.lo^T + lo.hi^T drops only
//                       the lo.lo^T term (~2^-22 relative to |f_i||f_j|), well inside the 1e-5 bar.
//                       Rows of all images are stacked; K is zero-padded to a multiple of 32.
//   k_lost_gram_tc      one CTA per 128x128 tile of one image, 192 threads, warp-specialised:
//                         warp 0 (one lane): TMA producer — per 32-wide K slab four bulk tensor loads
//                                  (A_hi, A_lo, B_hi, B_lo; 128 rows x 128 B, SWIZZLE_128B) into a
//                                  3-stage shared-memory ring, completion on mbarriers
//                         warp 1 (one lane): MMA issuer — per slab 4 k-steps x 3 tcgen05.mma
//                                  (hi.hi, hi.lo, lo.hi), M=128 N=128 K=8, fp32 accumulator in 128
//                                  TMEM columns; tcgen05.commit frees the stage / signals the epilogue
//                         warps 2-5: epilogue — tcgen05.ld (32 lanes x 32 columns per warp and step),
//                                  each thread owns one row: writes it to A and counts
//                                  (i != j ? max(A_ij, 0) : 0) > threshold, one atomic per row.
//   k_lost_gram_tc2<DIRECT>  the production kernel: CTA pairs (cta_group::2, 256x256 tiles, last column tile
//                       trimmed to a multiple of 16), persistent over a precomputed tile table; DIRECT reads the
//                       caller's features in place and derives the lo tiles in shared memory (no split kernel).
// Rows past an image's last patch read the next image's rows (or TMA zero fill at the very end):
// their products are computed and discarded by the epilogue's bounds checks.
#include "common.cuh"
#include "lost_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace b200p {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;                 // 16 KB: one operand tile of one stage
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;                // A_hi, A_lo, B_hi, B_lo
constexpr int TC_THREADS = 192;
constexpr size_t TC_SMEM_BYTES = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout_type [61,64) with SWIZZLE_128B = 2).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;                       // leading byte offset: 1 (unused by swizzled K-major layouts)
    d |= (uint64_t)(1024u >> 4) << 32;            // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, both K-major, N, M
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

// ---- split ---------------------------------------------------------------------------------------
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// one thread per float4 of the padded row
__global__ void __launch_bounds__(256)
k_lost_split_tf32(const float* __restrict__ feats, long long row_stride, int d, int d_pad,
                  const LostImageDev* __restrict__ meta, float* __restrict__ hi, float* __restrict__ lo, int vec_ok) {
    const LostImageDev im = meta[blockIdx.y];
    const int quads = d_pad >> 2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)im.n * quads) return;
    const int row = (int)(idx / quads), k0 = (int)(idx % quads) * 4;
    const float* p = feats + im.feat_off + (long long)row * row_stride + k0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec_ok && k0 + 3 < d) v = *reinterpret_cast<const float4*>(p);
    else {
        if (k0 + 0 < d) v.x = p[0];
        if (k0 + 1 < d) v.y = p[1];
        if (k0 + 2 < d) v.z = p[2];
        if (k0 + 3 < d) v.w = p[3];
    }
    float4 h, l;
    h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
    l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
    const long long o = ((long long)im.row_base + row) * d_pad + k0;
    *reinterpret_cast<float4*>(hi + o) = h;
    *reinterpret_cast<float4*>(lo + o) = l;
}

// ---- Gram on tcgen05: persistent, symmetric, double-buffered accumulator ---------------------------
// Tile schedule: only the upper-triangular 128x128 tiles (ti <= tj) of every image are computed;
// an off-diagonal tile is written twice (as is and transposed) and feeds the degree of its rows
// AND of its columns.  Each CTA (one per SM) walks tiles blockIdx.x, +gridDim.x, ... ; the smem ring
// and the two TMEM accumulators run across tile boundaries, so the epilogue of tile t overlaps the
// TMA/MMA main loop of tile t+1.
struct TileCoord { int img, ti, tj; };

__device__ __forceinline__ int find_image_pairs(const LostImageDev* __restrict__ meta, int n_images, int t) {
    int lo = 0, hi = n_images - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (meta[mid].pair_base <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ TileCoord decode_tile(const LostImageDev* __restrict__ meta, int n_images, int t) {
    TileCoord tc;
    tc.img = find_image_pairs(meta, n_images, t);
    const int T = meta[tc.img].tiles;
    int p = t - meta[tc.img].pair_base, ti = 0;
    while (p >= T - ti) { p -= T - ti; ++ti; }
    tc.ti = ti; tc.tj = ti + p;
    return tc;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

constexpr int TC_ACC = 2;                                       // TMEM accumulator buffers
constexpr int TC_TMEM_COLS2 = TC_ACC * TC_BN;                   // 256 columns

__global__ void __launch_bounds__(TC_THREADS, 1)
k_lost_gram_tc(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               const LostImageDev* __restrict__ meta, int n_images, int n_tiles, float* __restrict__ A_base,
               int* __restrict__ degree_base, float threshold, int d_pad) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B atoms need 1024-B alignment
    const uint32_t bar_base = smem_base + TC_STAGES * TC_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + TC_ACC + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * TC_STAGES + 2 * TC_ACC);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = d_pad / TC_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_hi); tma_prefetch_desc(&tm_lo);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < TC_ACC; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();                                  // reconverge after the one-lane setup above (.sync.aligned below)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"((uint32_t)TC_TMEM_COLS2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        int it = 0;                                                    // k-block counter across tiles
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const TileCoord tc = decode_tile(meta, n_images, t);
            const int a_row = meta[tc.img].row_base + tc.ti * TC_BM, b_row = meta[tc.img].row_base + tc.tj * TC_BN;
            const bool diag = tc.ti == tc.tj;
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);                     // slot free (passes at once the first time round)
                const uint32_t st = smem_base + s * TC_STAGE_BYTES;
                mbar_expect_tx(full_bar(s), diag ? 2 * TC_TILE_BYTES : 4 * TC_TILE_BYTES);
                tma_load_2d(st + 0 * TC_TILE_BYTES, &tm_hi, kb * TC_BK, a_row, full_bar(s));
                tma_load_2d(st + 1 * TC_TILE_BYTES, &tm_lo, kb * TC_BK, a_row, full_bar(s));
                if (!diag) {                                             // a diagonal tile multiplies the A tiles by themselves
                    tma_load_2d(st + 2 * TC_TILE_BYTES, &tm_hi, kb * TC_BK, b_row, full_bar(s));
                    tma_load_2d(st + 3 * TC_TILE_BYTES, &tm_lo, kb * TC_BK, b_row, full_bar(s));
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        const uint32_t idesc = umma_idesc_tf32(TC_BM, TC_BN);
        int it = 0, tl = 0;                                            // k-block / local tile counters
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tl) {
            const TileCoord tc = decode_tile(meta, n_images, t);
            const bool diag = tc.ti == tc.tj;
            const int acc = tl % TC_ACC;
            const uint32_t acc_ph = (uint32_t)(tl / TC_ACC) & 1u;
            mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u);               // epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_BN);
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                mbar_wait(full_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_base + s * TC_STAGE_BYTES;
                const uint64_t a_hi = umma_desc_sw128(st + 0 * TC_TILE_BYTES), a_lo = umma_desc_sw128(st + 1 * TC_TILE_BYTES);
                const uint64_t b_hi = diag ? a_hi : umma_desc_sw128(st + 2 * TC_TILE_BYTES);
                const uint64_t b_lo = diag ? a_lo : umma_desc_sw128(st + 3 * TC_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);    // 32 bytes per K=8 step inside the swizzle atom
                    umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                    umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                }
                umma_commit(empty_bar(s));                             // stage reusable once these MMAs have read it
            }
            umma_commit(tmem_full_bar(acc));                           // accumulator complete
        }
    } else if (warp >= 2) {
        // ===== epilogue: TMEM -> registers -> A (both triangles) + degrees =====
        const int q = warp & 3;                                         // TMEM lane quadrant this warp may access
        int tl = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tl) {
            const TileCoord tc = decode_tile(meta, n_images, t);
            const LostImageDev im = meta[tc.img];
            const int acc = tl % TC_ACC;
            const uint32_t acc_ph = (uint32_t)(tl / TC_ACC) & 1u;
            mbar_wait(tmem_full_bar(acc), acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row0 = tc.ti * TC_BM, col0 = tc.tj * TC_BN;
            const bool mirror = tc.ti != tc.tj;
            const int gi = row0 + q * 32 + lane;
            float* __restrict__ A = A_base + im.a_off;
            int* __restrict__ deg = degree_base + im.out_off;
            const bool vec_store = (im.n & 3) == 0 && (((uintptr_t)A) & 15u) == 0;
            int cnt = 0;
#pragma unroll 1
            for (int ch = 0; ch < TC_BN / 32; ++ch) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TC_BN + ch * 32), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int gj0 = col0 + ch * 32;
                int colcnt = 0;                                          // lane c ends up with the count of column gj0 + c
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int gj = gj0 + c;
                    const float v = __uint_as_float(r[c]);
                    const bool in = gi < im.n && gj < im.n;
                    const bool pos = in && ((gi == gj) ? 0.f : fmaxf(v, 0.f)) > threshold;
                    cnt += pos ? 1 : 0;
                    if (mirror) {
                        const unsigned bal = __ballot_sync(0xFFFFFFFFu, pos);
                        if (lane == c) colcnt = __popc(bal);
                        if (in) A[(long long)gj * im.n + gi] = v;         // transposed: lanes = consecutive addresses
                    }
                }
                if (mirror && colcnt && gj0 + lane < im.n) atomicAdd(deg + gj0 + lane, colcnt);
                if (gi < im.n) {
                    float* dst = A + (long long)gi * im.n + gj0;
                    if (vec_store && gj0 + 31 < im.n) {
#pragma unroll
                        for (int c = 0; c < 32; c += 4)
                            *reinterpret_cast<float4*>(dst + c) = make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]),
                                                                              __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) if (gj0 + c < im.n) dst[c] = __uint_as_float(r[c]);
                    }
                }
            }
            if (gi < im.n && cnt) atomicAdd(deg + gi, cnt);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tmem_empty_bar(acc));                          // 128 arrivals free the accumulator
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS2) : "memory");
    }
}

// ---- Gram on tcgen05, CTA pairs (cta_group::2) ----------------------------------------------------
// The single-CTA kernel above is bound by operand traffic: a 128x128 tile needs 64 KB of hi/lo operands
// per 768 tensor cycles.  Two CTAs of a cluster (adjacent SMs) compute one 256x256 tile together: CTA r
// supplies A rows [128r, +128) and the B half [128r, +128) of the tile's columns, the pair's tensor cores
// read both halves of B, and each CTA accumulates its 128 rows x 256 columns in its own TMEM — the same
// 64 KB per CTA and k-block now feed twice the flops.  Protocol:
//   * TMA loads in both CTAs use the .cta_group::2 form and report their bytes to the LEADER's (rank 0) full
//     barrier, which expects the bytes of both CTAs;
//   * the leader's elected thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 8) and commits with
//     .multicast::cluster to the empty / tmem_full barriers of BOTH CTAs;
//   * the epilogue warps of both CTAs arrive on the leader's tmem_empty barrier (256 arrivals);
//   * tile schedule: upper-triangular 256x256 tiles, one cluster walks tiles cluster_id, +n_clusters, ...
constexpr int T2_BM = 128, T2_BN = 256, T2_TILE = 256;           // rows per CTA, columns per tile, tile edge
constexpr int T2_ACC = 2, T2_TMEM_COLS = T2_ACC * T2_BN;         // 512 columns: all of TMEM
constexpr int T2_THREADS = 320;                                  // TMA warp, MMA warp, 8 epilogue warps
// direct mode: + converter warps deriving the lo tiles in shared memory: 4 next to the storing epilogue (168 registers per
// thread), 8 in count-only mode (79 registers): two per scheduler halve the time from "raw tile landed" to "lo tile ready"
__host__ __device__ constexpr int t2_threads(bool direct, int conv_warps) { return T2_THREADS + (direct ? 32 * conv_warps : 0); }
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;                      // clears the CTA-rank bit of a shared::cluster address (rank 0 of the pair)

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {       // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
// Arrival with the default semantics (release at CTA scope), the form CUTLASS's ClusterBarrier::arrive(cta_id) uses to
// hand smem written by transform warps (after fence.proxy.async) to a UMMA issued by the pair's leader.  The
// mbarrier.arrive.release.cluster form costs a MEMBAR.ALL.GPU per arrival.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(bar) : "memory");
}
// Same arrival without cluster-scope release of this thread's global stores: for barriers that only order
// tcgen05 work (the epilogue handing a TMEM accumulator back; tcgen05.fence::before_thread_sync does the
// ordering).  The release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR, i.e. every epilogue thread
// waited for its A stores to reach L2 before freeing the accumulator.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" :: "r"(bar) : "memory");
}

// columns the tensor cores compute for column tile tj of an image with n patches: the last tile is
// trimmed to the next multiple of 16 (the N granularity of cta_group::2), e.g. 144 instead of 256 at n = 900
__device__ __forceinline__ int tile2_cols(int n, int tj) {
    const int rest = (n - tj * T2_TILE + 15) & ~15;
    return rest < T2_BN ? rest : T2_BN;
}

__device__ __forceinline__ int find_image_pairs2(const LostImageDev* __restrict__ meta, int n_images, int t) {
    int lo = 0, hi = n_images - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (meta[mid].pair2_base <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ TileCoord decode_tile2(const LostImageDev* __restrict__ meta, int n_images, int t) {
    TileCoord tc;
    tc.img = find_image_pairs2(meta, n_images, t);
    const int T = (meta[tc.img].n + T2_TILE - 1) / T2_TILE;
    int p = t - meta[tc.img].pair2_base, ti = 0;
    while (p >= T - ti) { p -= T - ti; ++ti; }
    tc.ti = ti; tc.tj = ti + p;
    return tc;
}

// One record per 256x256 tile, built by k_lost_tile_table before the Gram kernel: the persistent roles read
// tile t + n_clusters while they work on tile t, so no role ever waits on the binary search over the image
// records (ten dependent L2 round trips per tile and role in the first version, ~20 % of the tile time in
// the TMA producer).
struct __align__(16) Tile2 {
    int a_row0, b_row0;        // first operand row of the tile's row / column block (before the rank offset)
    int info;                  // ncols | share << 9 | diag << 10 | ti << 12 | tj << 16 | K segment << 20 | last segment << 24
    int n;                     // patches of the image
    long long a_off, out_off;  // the image's offsets into A_base / degree_base
};
constexpr int kT2Share = 1 << 9, kT2Diag = 1 << 10, kT2Last = 1 << 24;
// K segmentation (keys wider than 512 when A is returned): the tensor core truncates every product to the accumulator's
// ulp, a bias of ~1.2e-8 d on same-sign sums (the squared norms on the diagonal: 2.2e-5 at d = 2048, bar 1e-5).  Segments
// of at most 12 k-blocks (384 keys) are accumulated from a zeroed TMEM accumulator each and added in fp32 by the
// epilogue (A in global memory carries the partial sum), which bounds the drift at the d = 384 level (4.9e-6).
// Count-only launches never segment: only the SIGN of off-diagonal entries matters there, and a mixed-sign sum near
// zero keeps a small accumulator (small ulp): its truncation error is ~1e-8 of |k_i||k_j| at any width.
constexpr int kSegBlocks = 12;

__global__ void __launch_bounds__(128)
k_lost_tile_table(const LostImageDev* __restrict__ meta, int n_images, int n_tiles, Tile2* __restrict__ tab, unsigned int* __restrict__ done,
                  int nseg) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (done && t < n_images) done[t] = 0u;          // count-only: per-image completion counters, see k_lost_finish
    if (done && t == 0) {                            // globaltimer trace in front of the counters: Gram start / end, finish start / end
        unsigned long long* tr = reinterpret_cast<unsigned long long*>(done) - 4;
        tr[0] = ~0ull; tr[1] = 0ull; tr[2] = ~0ull; tr[3] = 0ull;
    }
    if (t >= n_tiles) return;
    const TileCoord tc = decode_tile2(meta, n_images, t);
    const LostImageDev im = meta[tc.img];
    const int ncols = tile2_cols(im.n, tc.tj);
    Tile2 e;
    e.a_row0 = im.row_base + tc.ti * T2_TILE;
    e.b_row0 = im.row_base + tc.tj * T2_TILE;
    const bool diag = tc.ti == tc.tj;
    e.info = ncols | (diag && ncols == T2_BN ? kT2Share : 0) | (diag ? kT2Diag : 0) | (tc.ti << 12) | (tc.tj << 16);
    e.n = im.n; e.a_off = done ? (long long)tc.img : im.a_off; e.out_off = im.out_off;      // count-only: no A, the slot carries the image index
    for (int sg = 0; sg < nseg; ++sg) {             // the segments of a tile are consecutive records: one cluster runs them back to back
        Tile2 r = e;
        r.info |= (sg << 20) | (sg == nseg - 1 ? kT2Last : 0);
        tab[(long long)t * nseg + sg] = r;
    }
}
__device__ __forceinline__ Tile2 load_tile2(const Tile2* __restrict__ tab, int t, int n_tiles) {
    Tile2 e;
    const int4* p = reinterpret_cast<const int4*>(tab + (t < n_tiles ? t : n_tiles - 1));
    const int4 u = __ldg(p), v = __ldg(p + 1);
    e.a_row0 = u.x; e.b_row0 = u.y; e.info = u.z; e.n = u.w;
    e.a_off = ((long long)(unsigned)v.x) | ((long long)v.y << 32);
    e.out_off = ((long long)(unsigned)v.z) | ((long long)v.w << 32);
    return e;
}

// DIRECT = true: the TMA reads the caller's fp32 features in place (one tensor map with the caller's row
// stride, e.g. the k slice of a qkv buffer) and only the RAW tiles travel; four extra warps split every landed
// tile in shared memory the way k_lost_split_tf32 does in HBM - hi = tf32(x) over the raw word, lo = x - hi
// into a second ring at the same swizzled position - publish both to the async proxy and arrive on the
// leader's conv barrier the MMA issuer waits on.  No split kernel, no hi/lo arrays in HBM, half the L2 -> SM
// operand traffic.
//
// Shared memory (per CTA, 1024-B aligned):
//   pre-split:  3 stages x [A_hi | A_lo | B_hi | B_lo]                         (16 KB tiles)      192 KB
//   direct:     3 raw stages x [A_raw | B_raw]  +  3 lo stages x [A_lo | B_lo]                    192 KB
//               (the conversion of k-blocks i+1, i+2 overlaps the MMAs of i; 4 + 2 stages measure the same)
//   epilogue:   8 warps x 32 x 32 floats, 16-B XOR swizzle: the row-major copy of A is transposed through
//               shared memory so that one store instruction writes four full 128-B lines           32 KB
//   count-only (WRITE_A = false): the caller did not ask for A (b200p_lost_batched with d_A == NULL).  Nothing is stored:
//               the epilogue only counts positive entries per row and column, the finish kernel gets A[seed, :] and
//               M from the keys (two skinny mat-vecs).  No staging buffer: 193 KB, which leaves room on the SM for a
//               CTA of the finish kernel of the previous sub-batch (b200p_lost_batched overlaps the two).
constexpr int T2_LO_STAGES = 3;
constexpr int T2_RING_BYTES = TC_STAGES * TC_STAGE_BYTES;        // == (3 raw + 3 lo) * 2 tiles
constexpr int T2_STAGING_BYTES = 8 * 32 * 32 * 4;               // == one raw stage (2 tiles)
constexpr size_t T2_SMEM_BYTES = (size_t)T2_RING_BYTES + T2_STAGING_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
static_assert(T2_RING_BYTES == (3 + T2_LO_STAGES) * 2 * TC_TILE_BYTES, "ring layouts must have the same size");
constexpr size_t T2_SMEM_COUNT_BYTES = (size_t)T2_RING_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
static_assert(T2_SMEM_BYTES <= 227 * 1024, "shared memory budget");

template <bool DIRECT, bool WRITE_A, int T2_CONV_WARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(t2_threads(DIRECT, T2_CONV_WARPS), 1)
k_lost_gram_tc2(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                const Tile2* __restrict__ tab, int n_tiles, float* __restrict__ A_base,
                int* __restrict__ degree_base, float threshold, int d_pad, unsigned int* __restrict__ done, int nseg) {
    extern __shared__ uint8_t smem_raw[];
    // a dependent kernel (the count-only finish) may be scheduled beside this grid as soon as all of its CTAs are running
    if (!WRITE_A) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!WRITE_A && done && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long*>(done) - 4, lost_globaltimer());
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t staging_base = smem_base + T2_RING_BYTES;
    const uint32_t bar_base = staging_base + (WRITE_A ? T2_STAGING_BYTES : 0);      // count-only: no staging buffer (T2_SMEM_COUNT_BYTES)
    constexpr int T2_RAW_STAGES = 3;
    constexpr int NFULL = DIRECT ? T2_RAW_STAGES : TC_STAGES;     // stages the TMA fills
    auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // [4]
    auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };                // [4]
    auto conv_bar = [&](int s) { return bar_base + 8u * (8 + s); };                 // [4] direct: lo tiles ready (leader's copy is used)
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (16 + a); };           // [2]
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (18 + a); };          // [2]
    const uint32_t tmem_slot = bar_base + 8u * 20;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    // operand tile addresses of k-block `it`
    auto hi_tiles = [&](int it) {
        if (!DIRECT) return smem_base + (uint32_t)(it % TC_STAGES) * TC_STAGE_BYTES;
        const int s = it % T2_RAW_STAGES;
        return s < 3 ? smem_base + (uint32_t)s * (2 * TC_TILE_BYTES) : staging_base;
    };
    auto lo_tiles = [&](int it) { return DIRECT ? smem_base + (uint32_t)(3 * 2 + (it % T2_LO_STAGES) * 2) * TC_TILE_BYTES
                                                : smem_base + (uint32_t)(it % TC_STAGES) * TC_STAGE_BYTES + 2 * TC_TILE_BYTES; };
    // each pair of tiles is [A | B]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                    // 0 = leader of the pair
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int num_kb = d_pad / TC_BK;
    // work items of this cluster: (tile, K segment), the segments of a tile back to back; item j -> record index
    const int my_tiles = cluster_id < n_tiles ? (n_tiles - cluster_id + n_clusters - 1) / n_clusters : 0;
    const int n_items = my_tiles * nseg;
    auto rec_of = [&](int j) { const int jj = j < n_items ? j : n_items - 1; return (cluster_id + (jj / nseg) * n_clusters) * nseg + jj % nseg; };
    const int n_recs = n_tiles * nseg;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_hi);
        if (!DIRECT) tma_prefetch_desc(&tm_lo);
        for (int s = 0; s < 4; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 4; ++s) {
            mbar_init(conv_bar(s), 2 * T2_CONV_WARPS);          // one arrival per converter warp of both CTAs
        }
        for (int a = 0; a < T2_ACC; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 2 * 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"((uint32_t)T2_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                          // barriers of both CTAs are initialised, TMEM is allocated
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer (both CTAs): own A rows and own half of the tile's B columns
        int it = 0;
        Tile2 nxt = load_tile2(tab, rec_of(0), n_recs);
        for (int j = 0; j < n_items; ++j) {
            const Tile2 e = nxt;
            nxt = load_tile2(tab, rec_of(j + 1), n_recs);
            const int ncols = e.info & 511;
            const int a_row = e.a_row0 + (int)rank * T2_BM;
            const int b_row = e.b_row0 + (int)rank * (ncols >> 1);
            const bool share = (e.info & kT2Share) != 0;           // full diagonal tile: the B halves are the A tiles
            const int kb0 = nseg > 1 ? ((e.info >> 20) & 15) * kSegBlocks : 0, kb1 = nseg > 1 ? min(num_kb, kb0 + kSegBlocks) : num_kb;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % NFULL;
                const uint32_t ph = (uint32_t)(it / NFULL) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                const uint32_t hi = hi_tiles(it);
                if (DIRECT) {                                    // raw tiles only, bytes reported to this CTA's own barrier
                    mbar_expect_tx(full_bar(s), share ? TC_TILE_BYTES : 2 * TC_TILE_BYTES);
                    tma_load_2d(hi, &tm_hi, kb * TC_BK, a_row, full_bar(s));
                    if (!share) tma_load_2d(hi + TC_TILE_BYTES, &tm_hi, kb * TC_BK, b_row, full_bar(s));
                } else {                                         // bytes of both CTAs reported to the leader's barrier
                    const uint32_t lo = lo_tiles(it);
                    const uint32_t lbar = full_bar(s) & kPeerMask;
                    if (rank == 0) mbar_expect_tx(full_bar(s), 2u * (share ? 2 * TC_TILE_BYTES : 4 * TC_TILE_BYTES));
                    tma_load_2d_2sm(hi, &tm_hi, kb * TC_BK, a_row, lbar);
                    tma_load_2d_2sm(lo, &tm_lo, kb * TC_BK, a_row, lbar);
                    if (!share) {
                        tma_load_2d_2sm(hi + TC_TILE_BYTES, &tm_hi, kb * TC_BK, b_row, lbar);
                        tma_load_2d_2sm(lo + TC_TILE_BYTES, &tm_lo, kb * TC_BK, b_row, lbar);
                    }
                }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ===== MMA issuer (leader CTA only) =====
        int it = 0, tl = 0;
        Tile2 nxt = load_tile2(tab, rec_of(0), n_recs);
        for (int j = 0; j < n_items; ++j, ++tl) {
            const Tile2 e = nxt;
            nxt = load_tile2(tab, rec_of(j + 1), n_recs);
            const int ncols = e.info & 511;
            const bool share = (e.info & kT2Share) != 0;
            const uint32_t idesc = umma_idesc_tf32(2 * T2_BM, ncols);
            const int kb0 = nseg > 1 ? ((e.info >> 20) & 15) * kSegBlocks : 0, kb1 = nseg > 1 ? min(num_kb, kb0 + kSegBlocks) : num_kb;
            const int acc = tl % T2_ACC;
            const uint32_t acc_ph = (uint32_t)(tl / T2_ACC) & 1u;
            mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * T2_BN);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                if (DIRECT) mbar_wait(conv_bar(it % T2_LO_STAGES), (uint32_t)(it / T2_LO_STAGES) & 1u);    // raw tiles landed, lo tiles derived, in both CTAs
                else mbar_wait(full_bar(it % TC_STAGES), (uint32_t)(it / TC_STAGES) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t hi = hi_tiles(it), lo = lo_tiles(it);
                const uint64_t a_hi = umma_desc_sw128(hi), a_lo = umma_desc_sw128(lo);
                const uint64_t b_hi = share ? a_hi : umma_desc_sw128(hi + TC_TILE_BYTES);
                const uint64_t b_lo = share ? a_lo : umma_desc_sw128(lo + TC_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
                    umma_tf32_2sm(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    umma_tf32_2sm(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                    umma_tf32_2sm(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                }
                umma_commit_2sm(empty_bar(it % NFULL));             // frees the stage in both CTAs (direct: raw and lo stage, see below)
            }
            umma_commit_2sm(tmem_full_bar(acc));                    // both epilogues may read their half
        }
    } else if (DIRECT && warp >= T2_THREADS / 32) {
        // ===== converter (both CTAs, 4 warps): raw tile -> (hi in place, lo) =====
        const int ct = threadIdx.x - T2_THREADS;                     // 0..127
        int it = 0;
        Tile2 nxt = load_tile2(tab, rec_of(0), n_recs);
        for (int j = 0; j < n_items; ++j) {
            const bool share = (nxt.info & kT2Share) != 0;
            const int kb0 = nseg > 1 ? ((nxt.info >> 20) & 15) * kSegBlocks : 0, kb1 = nseg > 1 ? min(num_kb, kb0 + kSegBlocks) : num_kb;
            nxt = load_tile2(tab, rec_of(j + 1), n_recs);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                mbar_wait(full_bar(it % T2_RAW_STAGES), (uint32_t)(it / T2_RAW_STAGES) & 1u);          // this CTA's raw tiles have landed
                // the lo stage is free once the MMAs of k-block it - 3 are done: the commit that freed THAT k-block's raw
                // stage says so (a second tcgen05.commit per k-block for a separate barrier costs ~0.1 ms per call)
                if (it >= T2_LO_STAGES) {
                    const int j = it - T2_LO_STAGES;
                    mbar_wait(empty_bar(j % T2_RAW_STAGES), (uint32_t)(j / T2_RAW_STAGES) & 1u);
                }
                // plain shared-memory pointers: the compiler batches the eight loads of a tile ahead of the math and the stores
                uint4* __restrict__ src = reinterpret_cast<uint4*>(smem_raw + (hi_tiles(it) - smem_u32(smem_raw))) + ct;
                uint4* __restrict__ dst = reinterpret_cast<uint4*>(smem_raw + (lo_tiles(it) - smem_u32(smem_raw))) + ct;
                constexpr int QUADS = TC_TILE_BYTES / (16 * 32 * T2_CONV_WARPS);      // 8 float4 per thread and tile
#pragma unroll 1
                for (int tile = 0; tile < (share ? 1 : 2); ++tile) {
                    uint4 x[QUADS];
#pragma unroll
                    for (int j = 0; j < QUADS; ++j) x[j] = src[(tile * QUADS + j) * (32 * T2_CONV_WARPS)];
#pragma unroll
                    for (int j = 0; j < QUADS; ++j) {
                        // hi = tf32(x), rounded to nearest, replaces the raw word in place; lo = x - hi, exact and SIGNED.
                        // Feeding the raw word (the tensor core truncates it) with lo = x - trunc(x) saves this store and was
                        // the first version, but then hi.hi, hi.lo and lo.hi are all biased the same way, and the tensor core's
                        // accumulator, which truncates every product to its own ulp, drifts three times as far on same-sign
                        // sums: 1.8e-5 |k|^2 on the diagonal at d = 768 against 6e-6 with a rounded hi (bar: 1e-5).
                        // (hi by integer round-half-up on the 13 dropped bits, lo = x - hi exact and signed: 3 ALU instructions per key)
                        uint4 h, l;
                        h.x = (x[j].x + 0x1000u) & 0xFFFFE000u; h.y = (x[j].y + 0x1000u) & 0xFFFFE000u;
                        h.z = (x[j].z + 0x1000u) & 0xFFFFE000u; h.w = (x[j].w + 0x1000u) & 0xFFFFE000u;
                        l.x = __float_as_uint(__uint_as_float(x[j].x) - __uint_as_float(h.x)); l.y = __float_as_uint(__uint_as_float(x[j].y) - __uint_as_float(h.y));
                        l.z = __float_as_uint(__uint_as_float(x[j].z) - __uint_as_float(h.z)); l.w = __float_as_uint(__uint_as_float(x[j].w) - __uint_as_float(h.w));
                        src[(tile * QUADS + j) * (32 * T2_CONV_WARPS)] = h;
                        dst[(tile * QUADS + j) * (32 * T2_CONV_WARPS)] = l;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor cores
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(conv_bar(it % T2_LO_STAGES) & kPeerMask);
            }
        }
    } else if (warp >= 2 && warp < T2_THREADS / 32) {
        // ===== epilogue (both CTAs): rows [256 ti + 128 rank, +128) x 256 columns of the tile =====
        // 8 warps: TMEM lane quadrant q = warp & 3 (rows), column half hsel (4 chunks of 32 columns each)
        const int q = warp & 3, hsel = (warp - 2) >> 2;
        const uint32_t stg = staging_base + (uint32_t)(warp - 2) * 4096u;      // this warp's 32x32 transpose buffer
        int tl = 0;
        Tile2 nxt = load_tile2(tab, rec_of(0), n_recs);
        for (int j = 0; j < n_items; ++j, ++tl) {
            const Tile2 im = nxt;
            nxt = load_tile2(tab, rec_of(j + 1), n_recs);
            const bool seg_first = ((im.info >> 20) & 15) == 0, seg_last = (im.info & kT2Last) != 0;      // K segments (WRITE_A, wide keys)
            const int acc = tl % T2_ACC;
            const uint32_t acc_ph = (uint32_t)(tl / T2_ACC) & 1u;
            mbar_wait(tmem_full_bar(acc), acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row0 = (im.info >> 12 & 15) * T2_TILE + (int)rank * T2_BM, col0 = (im.info >> 16 & 15) * T2_TILE;
            const bool mirror = (im.info & kT2Diag) == 0;       // a diagonal 256-tile holds both triangles already
            const int gi0 = row0 + q * 32, gi = gi0 + lane;
            float* __restrict__ A = WRITE_A ? A_base + im.a_off : nullptr;
            int* __restrict__ deg = degree_base + im.out_off;
            const bool vec_store = !WRITE_A || ((im.n & 3) == 0 && (((uintptr_t)A) & 15u) == 0);      // count-only: the fast paths store nothing
            const float thr0 = fmaxf(threshold, 0.f);                  // (i != j ? max(A_ij, 0) : 0) > threshold  <=>  A_ij > thr0 off the diagonal
            int cnt = 0;
#pragma unroll 1
            for (int ch = 4 * hsel; ch < 4 * hsel + 4; ++ch) {
                const int gj0 = col0 + ch * 32;
                if (gj0 >= im.n) break;                          // warp-uniform: columns past the image
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * T2_BN + ch * 32), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (WRITE_A && !seg_first && gi < im.n) {
                    // later K segment: add the partial sum the earlier segments left in A (this lane's row, fp32 adds)
                    const float* __restrict__ prow = A + (long long)gi * im.n + gj0;
                    if (vec_store && gj0 + 32 <= im.n) {
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const float4 pv4 = *reinterpret_cast<const float4*>(prow + c);
                            r[c] = __float_as_uint(__uint_as_float(r[c]) + pv4.x); r[c + 1] = __float_as_uint(__uint_as_float(r[c + 1]) + pv4.y);
                            r[c + 2] = __float_as_uint(__uint_as_float(r[c + 2]) + pv4.z); r[c + 3] = __float_as_uint(__uint_as_float(r[c + 3]) + pv4.w);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) if (gj0 + c < im.n) r[c] = __float_as_uint(__uint_as_float(r[c]) + prow[c]);
                    }
                }
                if (WRITE_A && !seg_first) __syncwarp();             // everybody has read its partial row before the staged stores overwrite it
                if (WRITE_A && !vec_store && gi < im.n) {            // unaligned A: scalar row stores
                    float* dst = A + (long long)gi * im.n + gj0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) if (gj0 + c < im.n) dst[c] = __uint_as_float(r[c]);
                }
                if (WRITE_A && vec_store) {
                    // row-major copy through shared memory: lane i parks its row (8 float4, 16-B groups XOR-swizzled by
                    // i & 7: conflict-free both ways), then lane l picks up columns 4 (l & 7) .. +3 of rows 4 p + (l >> 3)
                    // and one store instruction covers four complete 128-B lines of A
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(stg + (uint32_t)lane * 128u + (uint32_t)((g ^ (lane & 7)) << 4)),
                                     "r"(r[4 * g]), "r"(r[4 * g + 1]), "r"(r[4 * g + 2]), "r"(r[4 * g + 3]) : "memory");
                    __syncwarp();
                    const int cg = lane & 7, rsub = lane >> 3;
                    const bool col_in = gj0 + 4 * cg < im.n;         // n % 4 == 0: a group of four columns is in or out as a whole
#pragma unroll
                    for (int pth = 0; pth < 8; ++pth) {
                        const int rr = 4 * pth + rsub;
                        uint32_t y0, y1, y2, y3;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(y0), "=r"(y1), "=r"(y2), "=r"(y3)
                                     : "r"(stg + (uint32_t)rr * 128u + (uint32_t)((cg ^ (rr & 7)) << 4)) : "memory");
                        if (col_in && gi0 + rr < im.n)
                            *reinterpret_cast<float4*>(A + (long long)(gi0 + rr) * im.n + gj0 + 4 * cg) =
                                make_float4(__uint_as_float(y0), __uint_as_float(y1), __uint_as_float(y2), __uint_as_float(y3));
                    }
                    __syncwarp();                                    // the buffer is rewritten by the next chunk
                }
                if (WRITE_A && !seg_last) continue;                  // partial sum parked in A; the last segment counts and mirrors
                // Fast paths, decided per 32x32 chunk: every element inside the image and none on the diagonal (a diagonal tile
                // holds the diagonal only in its chunks with gi0 == gj0).  The epilogue is bound by dependent ALU latency with two
                // warps per scheduler (clock64: 3100 cycles for a diagonal-tile chunk on the general path, without a single store),
                // so only the ragged last chunks and the 32 diagonal chunks of an image pay for the per-element tests.
                const bool full = vec_store && gi0 + 32 <= im.n && gj0 + 32 <= im.n && (mirror || gj0 != gi0);
                if (full && mirror) {
                    int colcnt = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                    float* __restrict__ col = WRITE_A ? A + (long long)gj0 * im.n + gi : nullptr;
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const bool p0 = __uint_as_float(r[c]) > thr0, p1 = __uint_as_float(r[c + 1]) > thr0;
                        const bool p2 = __uint_as_float(r[c + 2]) > thr0, p3 = __uint_as_float(r[c + 3]) > thr0;
                        c0 += p0; c1 += p1; c2 += p2; c3 += p3;
                        const unsigned b0 = __ballot_sync(0xFFFFFFFFu, p0), b1 = __ballot_sync(0xFFFFFFFFu, p1);
                        const unsigned b2 = __ballot_sync(0xFFFFFFFFu, p2), b3 = __ballot_sync(0xFFFFFFFFu, p3);
                        if ((lane >> 2) == (c >> 2)) colcnt = __popc((lane & 3) == 0 ? b0 : (lane & 3) == 1 ? b1 : (lane & 3) == 2 ? b2 : b3);
                        if (WRITE_A) {
                            col[(long long)(c + 0) * im.n] = __uint_as_float(r[c]);          // transposed: lanes = consecutive addresses
                            col[(long long)(c + 1) * im.n] = __uint_as_float(r[c + 1]);
                            col[(long long)(c + 2) * im.n] = __uint_as_float(r[c + 2]);
                            col[(long long)(c + 3) * im.n] = __uint_as_float(r[c + 3]);
                        }
                    }
                    cnt += (c0 + c1) + (c2 + c3);
                    if (colcnt) atomicAdd(deg + gj0 + lane, colcnt);
                    continue;
                }
                if (full) {                                       // diagonal tile, chunk off the diagonal: both triangles are computed, count only
                    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        c0 += __uint_as_float(r[c]) > thr0; c1 += __uint_as_float(r[c + 1]) > thr0;
                        c2 += __uint_as_float(r[c + 2]) > thr0; c3 += __uint_as_float(r[c + 3]) > thr0;
                    }
                    cnt += (c0 + c1) + (c2 + c3);
                    continue;
                }
                int colcnt = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int gj = gj0 + c;
                    const float v = __uint_as_float(r[c]);
                    const bool in = gi < im.n && gj < im.n;
                    const bool pos = in && ((gi == gj) ? 0.f : fmaxf(v, 0.f)) > threshold;
                    cnt += pos ? 1 : 0;
                    if (mirror) {
                        const unsigned bal = __ballot_sync(0xFFFFFFFFu, pos);
                        if (lane == c) colcnt = __popc(bal);
                        if (WRITE_A && in) A[(long long)gj * im.n + gi] = v;
                    }
                }
                if (mirror && colcnt && gj0 + lane < im.n) atomicAdd(deg + gj0 + lane, colcnt);
            }
            if (gi < im.n && cnt) atomicAdd(deg + gi, cnt);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive_cluster_relaxed(tmem_empty_bar(acc) & kPeerMask);  // 2 x 256 arrivals on the leader's barrier
            if (!WRITE_A && done) {
                // this warp's degree contributions of the tile are out: release them to the finish kernel (one count per warp)
                __syncwarp();
                if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(done + im.a_off) : "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                          // both CTAs are done with TMEM and with each other's barriers
    if (!WRITE_A && done && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(done) - 3, lost_globaltimer());
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)T2_TMEM_COLS) : "memory");
    }
}

// ---- host ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int make_map(CUtensorMap* map, const float* base, long long rows, int cols, long long row_stride) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("lost_batched: cuTensorMapEncodeTiled is not available from the driver"); return B200P_ECUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("lost_batched: cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")"); return B200P_ECUDA; }
    return B200P_OK;
}

// upper bound of the 256x256 tile records of a batch: an image of n <= 4096 patches has T = ceil(n / 256) <= 16
// tile rows and T (T + 1) / 2 <= 8.5 T tiles
static size_t tile_table_bytes(int n_images, long long total_patches) {
    return ((size_t)(9 * (total_patches / T2_TILE + n_images)) + 16) * sizeof(Tile2) * 16 /* K segments */ + 32 + (size_t)n_images * sizeof(unsigned int);
}

size_t lost_tc_workspace_bytes(int n_images, long long total_patches, int d) {
    const int d_pad = (d + TC_BK - 1) / TC_BK * TC_BK;
    return 2 * ((size_t)total_patches * d_pad * sizeof(float) + 1024) + tile_table_bytes(n_images, total_patches) + 256;
}

// The caller's features can be read in place by one tensor map when they form a 2-D array with a uniform,
// 16-byte-granular row stride in which every image starts on a row boundary.
bool lost_tc_direct_ok(const float* d_feats, long long row_stride, int d, const b200p_lost_image_t* h_meta, int n_images) {
    if ((((uintptr_t)d_feats) & 15u) != 0 || (row_stride & 3) != 0 || d < TC_BK || row_stride < d) return false;
    for (int b = 0; b < n_images; ++b) {
        if (h_meta[b].feat_offset % row_stride != 0) return false;
        if (h_meta[b].feat_offset / row_stride + (long long)h_meta[b].dim0 * h_meta[b].dim1 >= (1ll << 31)) return false;
    }
    return true;
}

// Set-up of one b200p_lost_batched call: tile table, tensor maps (and the hi/lo split for the pre-split modes).
int lost_gram_prepare(LostGramPlan* gp, const float* d_feats, long long row_stride, int d, const LostImageDev* d_meta,
                      const std::vector<LostImageDev>& meta, long long total_patches, int n_max, void* ws, size_t ws_bytes,
                      int vec_ok, cudaStream_t st, int mode, bool count_only) {
    static_assert(sizeof(CUtensorMap) == sizeof(gp->tm_hi), "tensor map storage");
    const int n_images = (int)meta.size();
    const int d_pad = (d + TC_BK - 1) / TC_BK * TC_BK;
    const LostImageDev& last = meta.back();
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    gp->mode = mode; gp->d_pad = d_pad; gp->sms = sms; gp->d_meta = d_meta; gp->n_images = n_images; gp->nseg = 1;
    CUtensorMap* tm_hi = reinterpret_cast<CUtensorMap*>(gp->tm_hi);
    CUtensorMap* tm_lo = reinterpret_cast<CUtensorMap*>(gp->tm_lo);
    const int T2 = (last.n + T2_TILE - 1) / T2_TILE;
    gp->n_tiles2 = last.pair2_base + T2 * (T2 + 1) / 2;
    gp->n_tiles1 = last.pair_base + last.tiles * (last.tiles + 1) / 2;
    // workspace: [tile table | hi | lo]
    const size_t tab_bytes = (tile_table_bytes(n_images, total_patches) + 255) / 256 * 256;
    Tile2* tab = (Tile2*)(((uintptr_t)ws + 15) & ~(uintptr_t)15);
    gp->tab = tab;
    gp->d_done = nullptr;
    if (mode != LOST_TC_SINGLE) {
        if (ws_bytes < tab_bytes) { set_error("lost_batched: tensor-core workspace too small"); return B200P_EINVAL; }
        const int work = gp->n_tiles2 > n_images ? gp->n_tiles2 : n_images;
        gp->nseg = 1;
        const int num_kb = d_pad / TC_BK;
        if (!count_only && num_kb > kSegBlocks) gp->nseg = (num_kb + kSegBlocks - 1) / kSegBlocks;
        if ((size_t)gp->n_tiles2 * gp->nseg * sizeof(Tile2) + 16 + 32 + (size_t)n_images * 4 > tab_bytes) { set_error("lost_batched: tensor-core workspace too small"); return B200P_EINVAL; }
        if (count_only) gp->d_done = reinterpret_cast<unsigned int*>(tab + (size_t)gp->n_tiles2 * gp->nseg) + 8;   // behind the tile records: 4 x u64 trace, counters
        k_lost_tile_table<<<(work + 127) / 128, 128, 0, st>>>(d_meta, n_images, gp->n_tiles2, tab, gp->d_done, gp->nseg);
        B200P_LAUNCH_CHECK("k_lost_tile_table");
    }
    ws = (char*)ws + tab_bytes; ws_bytes -= ws_bytes < tab_bytes ? ws_bytes : tab_bytes;
    if (mode == LOST_TC_PAIR_DIRECT) {
        // meta[].row_base holds each image's first row in the caller's array (set by the caller of this function)
        long long rows = 0;
        for (const LostImageDev& m : meta) if ((long long)m.row_base + m.n > rows) rows = (long long)m.row_base + m.n;
        int rc = make_map(tm_hi, d_feats, rows, d, row_stride); if (rc) return rc;
        *tm_lo = *tm_hi;
        return B200P_OK;
    }
    const size_t arr = ((size_t)total_patches * d_pad * sizeof(float) + 1023) / 1024 * 1024;
    if (ws_bytes < 2 * arr) { set_error("lost_batched: tensor-core workspace too small"); return B200P_EINVAL; }
    float* hi = (float*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    float* lo = (float*)((char*)hi + arr);
    if ((char*)lo + (size_t)total_patches * d_pad * sizeof(float) > (char*)ws + ws_bytes) {
        set_error("lost_batched: tensor-core workspace too small"); return B200P_EINVAL;
    }
    const long long quads = (long long)n_max * (d_pad / 4);
    dim3 sgrid((unsigned)((quads + 255) / 256), (unsigned)n_images);
    k_lost_split_tf32<<<sgrid, 256, 0, st>>>(d_feats, row_stride, d, d_pad, d_meta, hi, lo, vec_ok);
    B200P_LAUNCH_CHECK("k_lost_split_tf32");
    int rc = make_map(tm_hi, hi, total_patches, d_pad, d_pad); if (rc) return rc;
    return make_map(tm_lo, lo, total_patches, d_pad, d_pad);
}

// Converter warps of the count-only direct kernel.  8 (default): two per scheduler halve the time from "raw tile landed" to
// "lo tile ready" — ncu: 414 us / 256 images, tensor pipe 84 %, against 439 us / 74 % with 4.  With 8 the kernel holds
// 576 threads x 80 registers, and a CTA of the finish kernel no longer fits beside it (the register file is split four
// ways between the schedulers); with 4 a 256-thread finish CTA does fit and runs under the Gram of the following
// images, but that co-run costs the Gram kernel ~25 us and leaves a ~75 us tail: 0.522 ms per call against ~0.50 ms for
// 8 warps + finish afterwards.  B200P_LOST_CONV_WARPS=4 selects the co-resident variant.
int lost_conv_warps() {
    static const int cw = [] { const char* e = getenv("B200P_LOST_CONV_WARPS"); return (e && atoi(e) == 4) ? 4 : 8; }();
    return cw;
}

// Gram + degrees of the 256x256 tiles [t_begin, t_end) of the table (pair modes; images own consecutive tiles, so a
// range of images is a range of tiles).  A_base == nullptr: count-only.  The single-CTA cross-check kernel always runs
// the whole batch.
int lost_gram_run(const LostGramPlan& gp, int t_begin, int t_end, float* A_base, int* d_degree, cudaStream_t st) {
    const CUtensorMap& tm_hi = *reinterpret_cast<const CUtensorMap*>(gp.tm_hi);
    const CUtensorMap& tm_lo = *reinterpret_cast<const CUtensorMap*>(gp.tm_lo);
    static bool attr_set = false;
    if (!attr_set) {
        B200P_CUDA(cudaFuncSetAttribute(k_lost_gram_tc2<true, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM_BYTES));
        B200P_CUDA(cudaFuncSetAttribute(k_lost_gram_tc2<true, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM_COUNT_BYTES));
        B200P_CUDA(cudaFuncSetAttribute(k_lost_gram_tc2<true, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM_COUNT_BYTES));
        B200P_CUDA(cudaFuncSetAttribute(k_lost_gram_tc2<true, false, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        B200P_CUDA(cudaFuncSetAttribute(k_lost_gram_tc2<false, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM_BYTES));
        B200P_CUDA(cudaFuncSetAttribute(k_lost_gram_tc2<false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM_COUNT_BYTES));
        B200P_CUDA(cudaFuncSetAttribute(k_lost_gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
        // count-only kernels: ask for the full 228 KB shared-memory carve-out (their own 194 KB would select the 196 KB
        // configuration), so that a CTA of the finish kernel finds room on the same SM
        B200P_CUDA(cudaFuncSetAttribute(k_lost_gram_tc2<true, false, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        B200P_CUDA(cudaFuncSetAttribute(k_lost_gram_tc2<false, false, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    if (gp.mode == LOST_TC_SINGLE) {
        const int grid = gp.n_tiles1 < gp.sms ? gp.n_tiles1 : gp.sms;      // persistent: one CTA per SM
        k_lost_gram_tc<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tm_hi, tm_lo, gp.d_meta, gp.n_images, gp.n_tiles1, A_base, d_degree, 0.0f, gp.d_pad);
        B200P_LAUNCH_CHECK("k_lost_gram_tc");
        return B200P_OK;
    }
    const int nt = t_end - t_begin;
    if (nt <= 0) return B200P_OK;
    const int grid2 = 2 * (nt < gp.sms / 2 ? nt : gp.sms / 2);             // one cluster (CTA pair) per two SMs
    const Tile2* tab = gp.tab + (size_t)t_begin * gp.nseg;
    if (gp.mode == LOST_TC_PAIR_DIRECT) {
        if (A_base) k_lost_gram_tc2<true, true, 4><<<grid2, t2_threads(true, 4), T2_SMEM_BYTES, st>>>(tm_hi, tm_hi, tab, nt, A_base, d_degree, 0.0f, gp.d_pad, nullptr, gp.nseg);
        else if (lost_conv_warps() == 4)
            k_lost_gram_tc2<true, false, 4><<<grid2, t2_threads(true, 4), T2_SMEM_COUNT_BYTES, st>>>(tm_hi, tm_hi, tab, nt, nullptr, d_degree, 0.0f, gp.d_pad, gp.d_done, 1);
        else
            k_lost_gram_tc2<true, false, 8><<<grid2, t2_threads(true, 8), T2_SMEM_COUNT_BYTES, st>>>(tm_hi, tm_hi, tab, nt, nullptr, d_degree, 0.0f, gp.d_pad, gp.d_done, 1);
        B200P_LAUNCH_CHECK("k_lost_gram_tc2<direct>");
    } else {
        if (A_base) k_lost_gram_tc2<false, true, 4><<<grid2, T2_THREADS, T2_SMEM_BYTES, st>>>(tm_hi, tm_lo, tab, nt, A_base, d_degree, 0.0f, gp.d_pad, nullptr, gp.nseg);
        else        k_lost_gram_tc2<false, false, 4><<<grid2, T2_THREADS, T2_SMEM_COUNT_BYTES, st>>>(tm_hi, tm_lo, tab, nt, nullptr, d_degree, 0.0f, gp.d_pad, gp.d_done, 1);
        B200P_LAUNCH_CHECK("k_lost_gram_tc2");
    }
    return B200P_OK;
}

}  // namespace b200p
