// lost.cu — LOST object discovery (object_discovery.py:23-134), batched over images.
//
//   K6  k_lost_gram_ffma   A_b = F_b F_b^T in fp32 (object_discovery.py:39) with the degree count of
//                          patch_scoring (:77-87) fused into the epilogue: A is written once and never
//                          re-read for clone / fill_diagonal / clamp / compare / sum.
//   K7  k_lost_finish      one CTA per image: seed = argmin degree (:57), the k_patches lowest-degree
//                          patches with lowest-index-first ties (:60, stable argsort), similars (:61),
//                          M = sum of the similar rows of A in sorted order (:62), 4-connected component
//                          of M > 0 that holds the seed (scipy.ndimage.label, :104-107), box (:114-128).
//   plus the two stand-alone mirrors used by the Python functions patch_scoring() and detect_box().
//
// Varlen batching: image b has n_b = dim0*dim1 patches; CTAs are assigned to (image, tile) pairs
// through a prefix table so one launch covers images of different sizes.
#include "common.cuh"
#include "lost_common.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace b200p {

constexpr int kLostMaxPatches = 4096;     // per image (ViT-S/8 at 480x480 = 3600)
constexpr int kFinThreads = 512;
constexpr int kLostMaxTensorCoreWidth = 6144;  // widest key the pair kernels take (16 K segments of 384): ViT-B 768, VGG16 512, ResNet-50 2048 all fit

// image records travel as kernel arguments (no pageable-memcpy stream sync, no staging buffer)
constexpr int kMetaPerLaunch = 256;
struct MetaPack { LostImageDev m[kMetaPerLaunch]; };
__global__ void k_lost_set_meta(LostImageDev* __restrict__ dst, MetaPack pack, int n) {
    if ((int)threadIdx.x < n) dst[threadIdx.x] = pack.m[threadIdx.x];
}

// Uniform batches (a [B, N, d] tensor: same dims, offsets in arithmetic progression) need no upload at all: record b
// is record 0 plus b times a constant step.  (Copying 256 records out of the kernel-parameter bank costs 13 us: the
// constant cache serialises the thread-dependent addresses.)
__global__ void k_lost_gen_meta(LostImageDev* __restrict__ dst, LostImageDev first, LostImageDev step, int n) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    LostImageDev m = first;
    m.feat_off += (long long)b * step.feat_off; m.a_off += (long long)b * step.a_off; m.out_off += (long long)b * step.out_off;
    m.tile_base += b * step.tile_base; m.row_base += b * step.row_base;
    m.pair_base += b * step.pair_base; m.pair2_base += b * step.pair2_base;
    dst[b] = m;
}
static bool lost_meta_uniform(const std::vector<LostImageDev>& meta, LostImageDev& step) {
    if (meta.size() < 2) return false;
    const LostImageDev& a = meta[0];
    const LostImageDev& b1 = meta[1];
    memset(&step, 0, sizeof(step));
    step.feat_off = b1.feat_off - a.feat_off; step.a_off = b1.a_off - a.a_off; step.out_off = b1.out_off - a.out_off;
    step.tile_base = b1.tile_base - a.tile_base; step.row_base = b1.row_base - a.row_base;
    step.pair_base = b1.pair_base - a.pair_base; step.pair2_base = b1.pair2_base - a.pair2_base;
    for (size_t i = 1; i < meta.size(); ++i) {
        const LostImageDev& m = meta[i];
        const long long k = (long long)i;
        if (m.n != a.n || m.dim0 != a.dim0 || m.dim1 != a.dim1 || m.img_h != a.img_h || m.img_w != a.img_w ||
            m.s0 != a.s0 || m.s1 != a.s1 || m.tiles != a.tiles) return false;
        if (m.feat_off != a.feat_off + k * step.feat_off || m.a_off != a.a_off + k * step.a_off ||
            m.out_off != a.out_off + k * step.out_off || m.tile_base != a.tile_base + k * step.tile_base ||
            m.row_base != a.row_base + k * step.row_base || m.pair_base != a.pair_base + k * step.pair_base ||
            m.pair2_base != a.pair2_base + k * step.pair2_base) return false;
    }
    return true;
}

// ---- K6: fp32 Gram + degree -----------------------------------------------------------------
constexpr int BM = 128, BK = 16, GT = 256;

// One CTA computes a 128x128 tile of A_b: rows [ti*128, +128) x cols [tj*128, +128).
// 256 threads, 8x8 accumulators per thread, K in steps of 16 through double-buffered shared memory.
__global__ void __launch_bounds__(GT)
k_lost_gram_ffma(const float* __restrict__ feats, long long row_stride, int d,
                 const LostImageDev* __restrict__ meta, int n_images, float* __restrict__ A_base,
                 int* __restrict__ degree_base, float threshold, int vec_ok) {
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Bs[2][BK][BM + 4];
    __shared__ int s_rowcnt[BM];
    const int b = find_image(meta, n_images, blockIdx.x);
    const LostImageDev im = meta[b];
    const int local = blockIdx.x - im.tile_base;
    const int ti = local / im.tiles, tj = local % im.tiles;
    const int row0 = ti * BM, col0 = tj * BM;
    const float* __restrict__ F = feats + im.feat_off;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    if (tid < BM) s_rowcnt[tid] = 0;

    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

    // loader: thread t handles rows (t>>2) and (t>>2)+64, k-quad (t&3)*4
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    auto load_tile = [&](int base_row, int k0, float (&dst)[BK][BM + 4]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lrow + 64 * h;
            const int gr = base_row + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < im.n) {
                const float* p = F + (long long)gr * row_stride + k0 + lk;
                if (vec_ok && k0 + lk + 3 < d) v = *reinterpret_cast<const float4*>(p);
                else {
                    if (k0 + lk + 0 < d) v.x = p[0];
                    if (k0 + lk + 1 < d) v.y = p[1];
                    if (k0 + lk + 2 < d) v.z = p[2];
                    if (k0 + lk + 3 < d) v.w = p[3];
                }
            }
            dst[lk + 0][r] = v.x; dst[lk + 1][r] = v.y; dst[lk + 2][r] = v.z; dst[lk + 3][r] = v.w;
        }
    };

    const int ksteps = (d + BK - 1) / BK;
    load_tile(row0, 0, As[0]);
    load_tile(col0, 0, Bs[0]);
    __syncthreads();
    for (int ks = 0; ks < ksteps; ++ks) {
        const int cur = ks & 1;
        if (ks + 1 < ksteps) {
            load_tile(row0, (ks + 1) * BK, As[cur ^ 1]);
            load_tile(col0, (ks + 1) * BK, Bs[cur ^ 1]);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
        }
        __syncthreads();
    }

    // epilogue: write A, count (i != j ? max(A,0) : 0) > threshold per row
    float* __restrict__ A = A_base + im.a_off;
    const bool vec_store = (im.n & 3) == 0 && (((uintptr_t)A) & 15u) == 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int lr = (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
        const int gi = row0 + lr;
        if (gi >= im.n) continue;
        int cnt = 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int gj0 = col0 + half * 64 + tx * 4;
            float v[4] = {acc[r][half * 4 + 0], acc[r][half * 4 + 1], acc[r][half * 4 + 2], acc[r][half * 4 + 3]};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gj = gj0 + q;
                if (gj < im.n) {
                    const float e = (gi == gj) ? 0.f : fmaxf(v[q], 0.f);
                    cnt += (e > threshold) ? 1 : 0;
                }
            }
            if (vec_store && gj0 + 3 < im.n) {
                *reinterpret_cast<float4*>(A + (long long)gi * im.n + gj0) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (gj0 + q < im.n) A[(long long)gi * im.n + gj0 + q] = v[q];
            }
        }
        if (cnt) atomicAdd(&s_rowcnt[lr], cnt);
    }
    __syncthreads();
    if (tid < BM && row0 + tid < im.n && s_rowcnt[tid]) atomicAdd(degree_base + im.out_off + row0 + tid, s_rowcnt[tid]);
}

// ---- block helpers ------------------------------------------------------------------------------
// exclusive scan of one int per thread over the CTA; returns the exclusive prefix, total in *total
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
    __syncthreads();                       // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? s_warp[lane] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= o) wi += u; }
        if (lane < nw) s_warp[lane] = wi - w;
        if (lane == nw - 1) s_warp[32] = wi;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[warp] + incl - v;
}

// Connected component (4-neighbourhood, scipy.ndimage.label's default structure) of `fg` holding
// `seed`, then its bounding box.  s_comp: one byte per cell.  Returns false if the seed is background.
__device__ bool component_box(const unsigned char* __restrict__ s_fg, unsigned char* __restrict__ s_comp, int n,
                              int dim0, int dim1, int seed, int* s_red, int (&box)[4]) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = tid; j < n; j += nt) s_comp[j] = (j == seed && s_fg[j]) ? 1 : 0;
    __syncthreads();
    if (!s_fg[seed]) return false;
    while (true) {
        int changed = 0;
        for (int j = tid; j < n; j += nt) {
            if (s_fg[j] && !s_comp[j]) {
                const int r = j / dim1, c = j - r * dim1;
                if ((r > 0 && s_comp[j - dim1]) || (r + 1 < dim0 && s_comp[j + dim1]) ||
                    (c > 0 && s_comp[j - 1]) || (c + 1 < dim1 && s_comp[j + 1])) { s_comp[j] = 1; changed = 1; }
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    if (tid == 0) { s_red[0] = dim0; s_red[1] = dim1; s_red[2] = -1; s_red[3] = -1; }
    __syncthreads();
    int rmin = dim0, cmin = dim1, rmax = -1, cmax = -1;
    for (int j = tid; j < n; j += nt) {
        if (s_comp[j]) {
            const int r = j / dim1, c = j - r * dim1;
            rmin = min(rmin, r); rmax = max(rmax, r); cmin = min(cmin, c); cmax = max(cmax, c);
        }
    }
    if (rmax >= 0) { atomicMin(&s_red[0], rmin); atomicMin(&s_red[1], cmin); atomicMax(&s_red[2], rmax); atomicMax(&s_red[3], cmax); }
    __syncthreads();
    box[0] = s_red[0]; box[1] = s_red[1]; box[2] = s_red[2]; box[3] = s_red[3];   // ymin, xmin, ymax, xmax (inclusive)
    return true;
}

// pred = [s1*xmin, s0*ymin, min(s1*(xmax+1), W), min(s0*(ymax+1), H)]   (object_discovery.py:116-128)
__device__ __forceinline__ void write_box(float* out, const int (&b)[4], float s0, float s1, int img_h, int img_w) {
    out[0] = s1 * (float)b[1];
    out[1] = s0 * (float)b[0];
    float x1 = s1 * (float)(b[3] + 1), y1 = s0 * (float)(b[2] + 1);
    if (img_w > 0) x1 = fminf(x1, (float)img_w);
    if (img_h > 0) y1 = fminf(y1, (float)img_h);
    out[2] = x1; out[3] = y1;
}

// ---- K7: seed, seed expansion, box --------------------------------------------------------------
// dot product of two key rows by one warp: lane l takes float4 l, l + 32, ... (scalar elements when the layout is not
// 16-byte granular), then a butterfly sum: the same order on every run.
__device__ __forceinline__ float warp_dot(const float* __restrict__ a, const float* __restrict__ b, int d, bool vec) {
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    if (vec) {
        const int d4 = d >> 2;
        for (int c = lane; c < d4; c += 32) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(a) + c), y = *(reinterpret_cast<const float4*>(b) + c);
            acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
        }
        for (int c = 4 * d4 + lane; c < d; c += 32) acc = fmaf(__ldg(a + c), b[c], acc);
    } else {
        for (int c = lane; c < d; c += 32) acc = fmaf(__ldg(a + c), b[c], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    return acc;
}

// measurement aid (b200p_lost_finish_trace): globaltimer stamps of the finish CTA of image `g_fin_trace_img` at its phase
// boundaries: [0] ready (degrees complete) [1] degrees + histogram + seed [2] cut-off [3] potentials listed
// [4] similars known [5] similars ordered [6] sum of the similar keys [7] M [8] component + box
__device__ unsigned long long g_fin_stamps[16];
__device__ int g_fin_trace_img = 0;
__device__ __forceinline__ void fin_stamp(int i) {
    if (threadIdx.x == 0 && (int)blockIdx.x == g_fin_trace_img) g_fin_stamps[i] = lost_globaltimer();      // block index within its launch
}

// FROM_KEYS = false: A was materialised (the caller asked for it): similars and M are read from it, rows added in the
//                    reference's order (object_discovery.py:61-62).
// FROM_KEYS = true:  count-only Gram, no A anywhere.  A[seed, p] = k_seed . k_p for the <= k_patches potentials and
//                    M = K (sum of the similar keys) — two skinny mat-vecs over the image's keys instead of 2 n^2
//                    floats written and 100 rows read back.  Same signs as the reference wherever an entry is decidable
//                    in fp32 (|M_j| above the rounding of a length-d dot product); M itself is not bit-identical.
template <bool FROM_KEYS>
__global__ void __launch_bounds__(kFinThreads)
k_lost_finish(const LostImageDev* __restrict__ meta, const float* __restrict__ A_base, const int* __restrict__ degree_base,
              int k_patches, int n_max, int* __restrict__ seed_out, float* __restrict__ box_out, int* __restrict__ status_out,
              float* __restrict__ M_out, const float* __restrict__ feats, long long row_stride, int d, int vec_ok,
              const unsigned int* __restrict__ done, int img_base) {
    // dynamic shared memory, sized for the largest image of the batch (n_max):
    //   int deg[n_max] | int hist[n_max+1 (+pad)] | int list[1024] | int sorted[1024] | u8 flag[n_max] | u8 comp[n_max]
    //   | FROM_KEYS: float vsum[d_pad4] | float kseed[d_pad4] | u8 simflag[1024]
    extern __shared__ __align__(16) unsigned char s_dyn[];
    int* s_deg = reinterpret_cast<int*>(s_dyn);
    int* s_hist = s_deg + n_max;
    int* s_list = s_hist + ((n_max + 1 + 3) & ~3);
    int* s_sorted = s_list + 1024;
    unsigned char* s_flag = reinterpret_cast<unsigned char*>(s_sorted + 1024);   // foreground of M
    unsigned char* s_comp = s_flag + n_max;
    const int d_pad4 = (d + 3) & ~3;
    float* s_vsum = reinterpret_cast<float*>(s_comp + n_max);                     // n_max is a multiple of 16
    float* s_kseed = s_vsum + d_pad4;
    unsigned char* s_simflag = reinterpret_cast<unsigned char*>(s_kseed + d_pad4);
    __shared__ int s_warp[33];
    __shared__ unsigned long long s_best;
    __shared__ int s_red[4];
    __shared__ int s_cut[3];

    const int img = img_base + (int)blockIdx.x;
    const LostImageDev im = meta[img];
    const int n = im.n, tid = threadIdx.x, nt = blockDim.x;
    if (FROM_KEYS && done) {
        // launched as a programmatic dependent of the count-only Gram kernel: wait until every epilogue warp of every tile
        // of this image has released its degree contributions (16 warps per 256x256 tile)
        if (tid == 0) {
            const unsigned t2 = (unsigned)(n + 255) / 256u, expected = 16u * (t2 * (t2 + 1u) / 2u);
            unsigned seen = 0;
            // bounded (~2 s): a protocol error must surface as status 2, never as a hung GPU
            for (unsigned spin = 0; spin < (1u << 23); ++spin) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done + img) : "memory");
                if (seen >= expected) break;
                __nanosleep(256);
            }
            s_red[0] = seen >= expected ? 1 : 0;
        }
        if (tid == 0) atomicMin(const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(done)) - 2, lost_globaltimer());
        __syncthreads();
        const bool ready = s_red[0] != 0;
        __syncthreads();
        if (!ready) {
            if (tid == 0) { seed_out[img] = -1; status_out[img] = 2; float* o = box_out + 4 * (long long)img; o[0] = o[1] = o[2] = o[3] = 0.f; }
            return;
        }
    }
    fin_stamp(0);
    const float* __restrict__ A = FROM_KEYS ? nullptr : A_base + im.a_off;
    const float* __restrict__ F = FROM_KEYS ? feats + im.feat_off : nullptr;
    const int* __restrict__ deg = degree_base + im.out_off;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;

    // degrees, seed = lowest degree, lowest index among equals (stable argsort, object_discovery.py:57,88)
    if (tid == 0) s_best = ~0ull;
    for (int j = tid; j <= n; j += nt) s_hist[j] = 0;
    __syncthreads();
    unsigned long long best = ~0ull;
    for (int j = tid; j < n; j += nt) {
        const int dg = __ldcg(deg + j);                     // L2: the counts may have been produced by a kernel still in flight
        s_deg[j] = dg;
        const unsigned long long key = ((unsigned long long)(unsigned)dg << 32) | (unsigned)j;
        best = key < best ? key : best;
        atomicAdd(&s_hist[min(dg, n)], 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long u = __shfl_xor_sync(0xFFFFFFFFu, best, o); best = u < best ? u : best; }
    if ((tid & 31) == 0) atomicMin(&s_best, best);
    __syncthreads();
    const int seed = (int)(s_best & 0xFFFFFFFFu);
    if (FROM_KEYS) for (int c = tid; c < d; c += nt) s_kseed[c] = F[(long long)seed * row_stride + c];

    fin_stamp(1);
    // cut-off degree D: the k lowest-degree patches are those with degree < D plus the first
    // `quota` (by index) of degree == D
    const int kk = min(k_patches, n);
    {
        // scan the degree histogram in slabs of blockDim
        int carry = 0;
        if (tid == 0) { s_cut[0] = n + 1; s_cut[1] = 0; }
        __syncthreads();
        for (int base = 0; base <= n; base += nt) {
            const int j = base + tid;
            const int v = j <= n ? s_hist[j] : 0;
            int total;
            const int excl = block_excl_scan(v, s_warp, &total) + carry;
            if (v > 0 && excl < kk && kk <= excl + v) { s_cut[0] = j; s_cut[1] = kk - excl; }   // unique j
            carry += total;
            __syncthreads();
        }
    }
    __syncthreads();
    const int D = s_cut[0], quota = s_cut[1];
    fin_stamp(2);
    if (FROM_KEYS) {
        // potentials in index order -> s_sorted (scratch), then one warp per potential: A[seed, p] = k_seed . k_p
        int carry = 0, n_pot = 0;
        for (int base = 0; base < n; base += nt) {
            const int j = base + tid;
            const int is_eq = (j < n && s_deg[j] == D) ? 1 : 0;
            int total;
            const int eq_rank = block_excl_scan(is_eq, s_warp, &total) + carry;
            carry += total;
            const int member = (j < n && (s_deg[j] < D || (is_eq && eq_rank < quota))) ? 1 : 0;
            int tot2;
            const int pos = block_excl_scan(member, s_warp, &tot2) + n_pot;
            if (member && pos < 1024) s_sorted[pos] = j;
            n_pot += tot2;
            __syncthreads();
        }
        n_pot = min(n_pot, 1024);
        fin_stamp(3);
        if (vec_ok && (d & 3) == 0) {
            // four potentials per warp and round: their loads are in flight together (the phase is pure latency)
            const int d4 = d >> 2;
            const float4* __restrict__ ks4 = reinterpret_cast<const float4*>(s_kseed);
            for (int i0 = 4 * warp; i0 < n_pot; i0 += 4 * nwarps) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                const float4* rowp[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) rowp[u] = reinterpret_cast<const float4*>(F + (long long)s_sorted[min(i0 + u, n_pot - 1)] * row_stride);
                for (int c = lane; c < d4; c += 32) {
                    float4 x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) x[u] = __ldg(rowp[u] + c);
                    const float4 y = ks4[c];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc[u] = fmaf(x[u].x, y.x, acc[u]); acc[u] = fmaf(x[u].y, y.y, acc[u]);
                        acc[u] = fmaf(x[u].z, y.z, acc[u]); acc[u] = fmaf(x[u].w, y.w, acc[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xFFFFFFFFu, acc[u], o);
                    if (lane == 0 && i0 + u < n_pot) s_simflag[i0 + u] = acc[u] > 0.0f ? 1 : 0;      // object_discovery.py:61 on the unmodified A
                }
            }
        } else {
            for (int i = warp; i < n_pot; i += nwarps) {
                const float a = warp_dot(F + (long long)s_sorted[i] * row_stride, s_kseed, d, vec_ok != 0);
                if (lane == 0) s_simflag[i] = a > 0.0f ? 1 : 0;          // object_discovery.py:61 on the unmodified A
            }
        }
        __syncthreads();
        int n_sim = 0;
        for (int base = 0; base < n_pot; base += nt) {
            const int i = base + tid;
            const int sim = (i < n_pot && s_simflag[i]) ? 1 : 0;
            int tot2;
            const int pos = block_excl_scan(sim, s_warp, &tot2) + n_sim;
            if (sim) s_list[pos] = s_sorted[i];
            n_sim += tot2;
            __syncthreads();
        }
        if (tid == 0) s_cut[2] = n_sim;
    } else {
        // membership + similars (A[seed, p] > 0 on the unmodified A, object_discovery.py:61)
        int carry = 0, n_sim = 0;
        for (int base = 0; base < n; base += nt) {
            const int j = base + tid;
            const int is_eq = (j < n && s_deg[j] == D) ? 1 : 0;
            int total;
            const int eq_rank = block_excl_scan(is_eq, s_warp, &total) + carry;
            carry += total;
            int sim = 0;
            if (j < n) {
                const bool member = s_deg[j] < D || (is_eq && eq_rank < quota);
                sim = (member && A[(long long)seed * n + j] > 0.0f) ? 1 : 0;
            }
            int tot2;
            const int pos = block_excl_scan(sim, s_warp, &tot2) + n_sim;
            if (sim && pos < 1024) s_list[pos] = j;
            n_sim += tot2;
            __syncthreads();
        }
        if (tid == 0) s_cut[2] = min(n_sim, 1024);
    }
    __syncthreads();
    const int n_sim = s_cut[2];
    fin_stamp(4);
    // order of the similars = order of `potentials` (ascending degree, index): rank by counting
    for (int i = tid; i < n_sim; i += nt) {
        const int p = s_list[i], dp = s_deg[p];
        int rank = 0;
        for (int q = 0; q < n_sim; ++q) {
            const int o = s_list[q], dq = s_deg[o];
            rank += (dq < dp || (dq == dp && o < p)) ? 1 : 0;
        }
        s_sorted[rank] = p;
    }
    __syncthreads();
    fin_stamp(5);
    if (FROM_KEYS) {
        // v = sum of the similar keys, added in that order; then M_j = k_j . v, one warp per row, four rows in flight
        for (int c = tid; c < d_pad4; c += nt) {
            float v = 0.f;
            if (c < d) {
                int r = 0;
                for (; r + 8 <= n_sim; r += 8) {
                    float x[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) x[u] = __ldg(F + (long long)s_sorted[r + u] * row_stride + c);
#pragma unroll
                    for (int u = 0; u < 8; ++u) v = __fadd_rn(v, x[u]);
                }
                for (; r < n_sim; ++r) v = __fadd_rn(v, __ldg(F + (long long)s_sorted[r] * row_stride + c));
            }
            s_vsum[c] = v;
        }
        __syncthreads();
        fin_stamp(6);
        if (vec_ok && (d & 3) == 0) {
            // one warp per group of ROWS rows: ROWS x (d / 128) independent 16-byte loads in flight per lane before the first FMA
            constexpr int ROWS = 4;        // 62 registers: a 256-thread CTA of this kernel fits beside a Gram CTA (6 rows: 86 registers, it no longer does)
            const int d4 = d >> 2;
            const float4* __restrict__ v4 = reinterpret_cast<const float4*>(s_vsum);
            for (int j0 = ROWS * warp; j0 < n; j0 += ROWS * nwarps) {
                float acc[ROWS];
                const float4* rowp[ROWS];
#pragma unroll
                for (int u = 0; u < ROWS; ++u) { acc[u] = 0.f; rowp[u] = reinterpret_cast<const float4*>(F + (long long)min(j0 + u, n - 1) * row_stride); }
                for (int c = lane; c < d4; c += 32) {
                    float4 x[ROWS];
#pragma unroll
                    for (int u = 0; u < ROWS; ++u) x[u] = __ldg(rowp[u] + c);
                    const float4 y = v4[c];
#pragma unroll
                    for (int u = 0; u < ROWS; ++u) {
                        acc[u] = fmaf(x[u].x, y.x, acc[u]); acc[u] = fmaf(x[u].y, y.y, acc[u]);
                        acc[u] = fmaf(x[u].z, y.z, acc[u]); acc[u] = fmaf(x[u].w, y.w, acc[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < ROWS; ++u) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xFFFFFFFFu, acc[u], o);
                    if (lane == 0 && j0 + u < n) {
                        s_flag[j0 + u] = acc[u] > 0.0f ? 1 : 0;
                        if (M_out) M_out[im.out_off + j0 + u] = acc[u];
                    }
                }
            }
        } else {
            for (int j = warp; j < n; j += nwarps) {
                const float m = warp_dot(F + (long long)j * row_stride, s_vsum, d, false);
                if (lane == 0) { s_flag[j] = m > 0.0f ? 1 : 0; if (M_out) M_out[im.out_off + j] = m; }
            }
        }
    } else {
        // M = sum over similars of A[s, :], rows added in that order (object_discovery.py:62)
        for (int j = tid; j < n; j += nt) {
            float m = 0.f;
            int r = 0;
            for (; r + 8 <= n_sim; r += 8) {                       // 8 independent loads in flight, then the ordered adds
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldg(A + (long long)s_sorted[r + u] * n + j);
#pragma unroll
                for (int u = 0; u < 8; ++u) m = __fadd_rn(m, v[u]);
            }
            for (; r < n_sim; ++r) m = __fadd_rn(m, __ldg(A + (long long)s_sorted[r] * n + j));
            s_flag[j] = m > 0.0f ? 1 : 0;
            if (M_out) M_out[im.out_off + j] = m;
        }
    }
    __syncthreads();
    fin_stamp(7);
    int box[4];
    const bool ok = component_box(s_flag, s_comp, n, im.dim0, im.dim1, seed, s_red, box);
    fin_stamp(8);
    if (tid == 0) {
        seed_out[img] = seed;
        status_out[img] = ok ? 0 : 1;
        float* o = box_out + 4 * (long long)img;
        if (ok) write_box(o, box, im.s0, im.s1, im.img_h, im.img_w);
        else { o[0] = o[1] = o[2] = o[3] = 0.f; }
        if (FROM_KEYS && done) atomicMax(const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(done)) - 1, lost_globaltimer());
    }
}

// ---- stand-alone mirrors -------------------------------------------------------------------------
// degree of a given matrix (patch_scoring, object_discovery.py:72-90): one warp per row
__global__ void __launch_bounds__(256)
k_lost_degree(const float* __restrict__ A, int n, long long lda, float threshold, int* __restrict__ degree) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    int cnt = 0;
    for (int j = lane; j < n; j += 32) {
        const float v = A[(long long)row * lda + j];
        const float e = (j == row) ? 0.f : fmaxf(v, 0.f);     // fill_diagonal_(0); A[A < 0] = 0
        cnt += (e > threshold) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if (lane == 0) degree[row] = cnt;
}
// sel[rank] = i where rank = #{j : (deg_j, j) < (deg_i, i)}: stable ascending-degree order
__global__ void __launch_bounds__(256)
k_lost_rank_by_degree(const int* __restrict__ degree, int n, long long* __restrict__ sel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int di = degree[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) { const int dj = __ldg(degree + j); rank += (dj < di || (dj == di && j < i)) ? 1 : 0; }
    sel[rank] = i;
}
__global__ void __launch_bounds__(kFinThreads)
k_lost_detect_box(const float* __restrict__ M, int n, int dim0, int dim1, int seed, float s0, float s1,
                  int img_h, int img_w, float* __restrict__ box_out, int* __restrict__ feat_box_out, int* __restrict__ status_out) {
    __shared__ unsigned char s_fg[kLostMaxPatches];
    __shared__ unsigned char s_comp[kLostMaxPatches];
    __shared__ int s_red[4];
    for (int j = threadIdx.x; j < n; j += blockDim.x) s_fg[j] = M[j] > 0.0f ? 1 : 0;
    __syncthreads();
    int box[4];
    const bool ok = component_box(s_fg, s_comp, n, dim0, dim1, seed, s_red, box);
    if (threadIdx.x == 0) {
        status_out[0] = ok ? 0 : 1;
        if (ok) {
            write_box(box_out, box, s0, s1, img_h, img_w);
            feat_box_out[0] = box[0]; feat_box_out[1] = box[1]; feat_box_out[2] = box[2] + 1; feat_box_out[3] = box[3] + 1;
        }
    }
}

}  // namespace b200p

using namespace b200p;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// The pair kernels (TC2 / TC2D on keys the tensor cores take) run count-only when the caller does not ask for A:
// no Gram matrix is materialised anywhere, the finish kernel works from the keys.
static bool lost_count_only(int gram_impl, int d) {
    return (gram_impl == B200P_LOST_GRAM_TC2 || gram_impl == B200P_LOST_GRAM_TC2D) && d <= kLostMaxTensorCoreWidth;
}

static const unsigned int* g_last_done = nullptr;
// measurement aid: globaltimer (ns) trace of the last count-only call on this thread's device: Gram first CTA start, Gram
// last CTA end, first finish CTA past its wait, last finish CTA end.  Synchronises the device.
extern "C" int b200p_lost_last_trace(uint64_t* h_out4) {
    B200P_REQUIRE(h_out4 != nullptr && g_last_done != nullptr, B200P_ESTATE, "lost_last_trace: no count-only call yet");
    B200P_CUDA(cudaDeviceSynchronize());
    B200P_CUDA(cudaMemcpy(h_out4, reinterpret_cast<const unsigned long long*>(g_last_done) - 4, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return B200P_OK;
}

extern "C" int b200p_lost_finish_trace(int image, uint64_t* h_out16) {
    B200P_REQUIRE(h_out16 != nullptr && image >= 0, B200P_EINVAL, "lost_finish_trace: bad argument");
    B200P_CUDA(cudaDeviceSynchronize());
    B200P_CUDA(cudaMemcpyFromSymbol(h_out16, g_fin_stamps, sizeof(unsigned long long) * 16));
    B200P_CUDA(cudaMemcpyToSymbol(g_fin_trace_img, &image, sizeof(int)));      // image traced by the NEXT call
    return B200P_OK;
}

extern "C" int b200p_lost_workspace_bytes(int n_images, int64_t total_patches, int64_t total_a, int d, int gram_impl, int64_t* out) {
    B200P_REQUIRE(out != nullptr && n_images >= 0 && total_patches >= 0 && total_a >= 0 && d >= 1, B200P_EINVAL, "lost_workspace_bytes: bad argument");
    if (lost_count_only(gram_impl, d)) total_a = 0;
    size_t bytes = align_up((size_t)n_images * sizeof(LostImageDev), 256) + align_up((size_t)total_a * sizeof(float), 256) + 256;
    if (gram_impl != B200P_LOST_GRAM_FFMA) bytes += align_up(lost_tc_workspace_bytes(n_images, total_patches, d), 256);
    *out = (int64_t)bytes;
    return B200P_OK;
}

extern "C" int b200p_lost_batched(int device, const float* d_feats, int64_t row_stride, int d,
                                  const b200p_lost_image_t* h_meta, int n_images, int k_patches,
                                  float* d_A, int32_t* d_degree, int32_t* d_seed, float* d_box,
                                  int32_t* d_status, void* d_workspace, int64_t workspace_bytes,
                                  int gram_impl, void* stream) {
    B200P_REQUIRE(d_feats && h_meta && d_degree && d_seed && d_box && d_status && d_workspace, B200P_EINVAL, "lost_batched: null argument");
    B200P_REQUIRE(n_images >= 1 && d >= 1 && row_stride >= d && k_patches >= 1, B200P_EINVAL, "lost_batched: bad sizes");
    B200P_REQUIRE(gram_impl >= B200P_LOST_GRAM_FFMA && gram_impl <= B200P_LOST_GRAM_TC2D, B200P_EINVAL, "lost_batched: bad gram_impl");
    B200P_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<LostImageDev> meta(n_images);
    long long tile_base = 0, pair_base = 0, pair2_base = 0, total_a = 0, total_patches = 0, deg_lo = -1, deg_hi = 0;
    int n_max = 0;
    B200P_REQUIRE(k_patches <= 1024, B200P_EINVAL, "lost_batched: k_patches must be <= 1024");
    bool vec = (((uintptr_t)d_feats) & 15u) == 0 && (row_stride & 3) == 0;
    const bool count_only = d_A == nullptr && lost_count_only(gram_impl, d);
    // The tensor cores truncate every product to the accumulator's ulp, so a same-sign sum (the squared norms on the
    // diagonal) drifts by ~1.2e-8 d |k|^2: 4.9e-6 at d = 384, 9.3e-6 at 768, 2.2e-5 at 2048 in ONE accumulation
    // (tools/gram_error_probe.py), whatever the operand split.  The pair kernels keep the 1e-5 bar for wide keys (ResNet-50
    // features, main_lost_original.py:277-280) by accumulating K segments of 384 from zero and adding them in fp32
    // (lost_tc.cu, kSegBlocks); only keys past 6144 go to the fp32 FMA Gram.  The single-CTA cross-check kernel (TC) does
    // not segment: an explicit request for it is honoured as is.
    if ((gram_impl == B200P_LOST_GRAM_TC2D || gram_impl == B200P_LOST_GRAM_TC2) && d > kLostMaxTensorCoreWidth) gram_impl = B200P_LOST_GRAM_FFMA;
    int tc_mode = gram_impl == B200P_LOST_GRAM_TC ? LOST_TC_SINGLE : LOST_TC_PAIR;
    if (gram_impl == B200P_LOST_GRAM_TC2D && lost_tc_direct_ok(d_feats, (long long)row_stride, d, h_meta, n_images)) tc_mode = LOST_TC_PAIR_DIRECT;
    for (int b = 0; b < n_images; ++b) {
        const b200p_lost_image_t& h = h_meta[b];
        const long long n = (long long)h.dim0 * h.dim1;
        B200P_REQUIRE(h.dim0 >= 1 && h.dim1 >= 1 && n <= kLostMaxPatches, B200P_EINVAL, "lost_batched: dims must give 1..4096 patches");
        B200P_REQUIRE(h.feat_offset >= 0 && h.out_offset >= 0 && h.a_offset >= 0, B200P_EINVAL, "lost_batched: negative offset");
        LostImageDev& m = meta[b];
        m.feat_off = h.feat_offset; m.out_off = h.out_offset; m.n = (int)n; m.dim0 = h.dim0; m.dim1 = h.dim1;
        m.img_h = h.img_h; m.img_w = h.img_w; m.s0 = h.scale0; m.s1 = h.scale1;
        m.a_off = d_A ? h.a_offset : total_a;
        m.tiles = (int)((n + BM - 1) / BM);
        m.tile_base = (int)tile_base;
        m.row_base = tc_mode == LOST_TC_PAIR_DIRECT ? (int)(h.feat_offset / row_stride) : (int)total_patches;
        m.pair_base = (int)pair_base;
        pair_base += (long long)m.tiles * (m.tiles + 1) / 2;
        m.pair2_base = (int)pair2_base; m.pad_ = 0;
        { const long long t2 = (n + 255) / 256; pair2_base += t2 * (t2 + 1) / 2; }
        tile_base += (long long)m.tiles * m.tiles;
        total_a += n * n;
        total_patches += n;
        if (h.feat_offset & 3) vec = false;
        if (deg_lo < 0 || h.out_offset < deg_lo) deg_lo = h.out_offset;
        if (h.out_offset + n > deg_hi) deg_hi = h.out_offset + n;
        if (n > n_max) n_max = (int)n;
    }
    B200P_REQUIRE(tile_base < (1ll << 31), B200P_EINVAL, "lost_batched: too many tiles in one call");
    const size_t meta_bytes = align_up((size_t)n_images * sizeof(LostImageDev), 256);
    const size_t a_bytes = (d_A || count_only) ? 0 : align_up((size_t)total_a * sizeof(float), 256);
    const size_t tc_bytes = gram_impl != B200P_LOST_GRAM_FFMA ? align_up(lost_tc_workspace_bytes(n_images, total_patches, d), 256) : 0;
    const size_t need = meta_bytes + a_bytes + tc_bytes;
    B200P_REQUIRE((size_t)workspace_bytes >= need, B200P_EINVAL, "lost_batched: workspace too small (see b200p_lost_workspace_bytes)");
    LostImageDev* d_meta = (LostImageDev*)d_workspace;
    float* A_base = d_A ? d_A : count_only ? nullptr : (float*)((char*)d_workspace + meta_bytes);
    LostImageDev step;
    if (lost_meta_uniform(meta, step)) {
        k_lost_gen_meta<<<(n_images + 255) / 256, 256, 0, st>>>(d_meta, meta[0], step, n_images);
        B200P_LAUNCH_CHECK("k_lost_gen_meta");
    } else
    for (int b0 = 0; b0 < n_images; b0 += kMetaPerLaunch) {
        MetaPack pack;
        const int nb = n_images - b0 < kMetaPerLaunch ? n_images - b0 : kMetaPerLaunch;
        for (int i = 0; i < nb; ++i) pack.m[i] = meta[b0 + i];
        k_lost_set_meta<<<1, kMetaPerLaunch, 0, st>>>(d_meta + b0, pack, nb);
        B200P_LAUNCH_CHECK("k_lost_set_meta");
    }
    // the call owns d_degree[min out_offset, max out_offset + n): cleared in one go
    B200P_CUDA(cudaMemsetAsync(d_degree + deg_lo, 0, (size_t)(deg_hi - deg_lo) * sizeof(int32_t), st));
    n_max = (n_max + 15) & ~15;
    size_t fin_smem = (size_t)n_max * 4 + (size_t)((n_max + 1 + 3) & ~3) * 4 + 2 * 1024 * 4 + 2 * (size_t)n_max;
    if (count_only) fin_smem += 2 * (size_t)((d + 3) & ~3) * 4 + 1024;
    B200P_REQUIRE(fin_smem <= 96 * 1024, B200P_EINVAL, "lost_batched: keys too wide for the finish kernel's shared memory");
    static bool attr_set = false;
    if (!attr_set) {
        B200P_CUDA(cudaFuncSetAttribute(k_lost_finish<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        B200P_CUDA(cudaFuncSetAttribute(k_lost_finish<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        B200P_CUDA(cudaFuncSetAttribute(k_lost_finish<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    if (gram_impl == B200P_LOST_GRAM_FFMA) {
        k_lost_gram_ffma<<<(int)tile_base, GT, 0, st>>>(d_feats, (long long)row_stride, d, d_meta, n_images, A_base, d_degree,
                                                       0.0f, vec ? 1 : 0);
        B200P_LAUNCH_CHECK("k_lost_gram_ffma");
    }
    LostGramPlan gp;
    if (gram_impl != B200P_LOST_GRAM_FFMA) {
        void* tc_ws = (char*)d_workspace + meta_bytes + a_bytes;
        int rc = lost_gram_prepare(&gp, d_feats, (long long)row_stride, d, d_meta, meta, total_patches, n_max, tc_ws, tc_bytes,
                                   vec ? 1 : 0, st, tc_mode, count_only);
        if (rc) return rc;
    }
    if (!count_only) {
        if (gram_impl != B200P_LOST_GRAM_FFMA) { int rc = lost_gram_run(gp, 0, gp.n_tiles2, A_base, d_degree, st); if (rc) return rc; }
        k_lost_finish<false><<<n_images, kFinThreads, fin_smem, st>>>(d_meta, A_base, d_degree, k_patches, n_max, d_seed, d_box, d_status, nullptr,
                                                                      d_feats, (long long)row_stride, d, vec ? 1 : 0, nullptr, 0);
        B200P_LAUNCH_CHECK("k_lost_finish");
        return B200P_OK;
    }
    // Count-only.  The finish kernel's cost is the mat-vec M = K v: it re-reads every key (355 MB per 256 images) and runs at
    // DRAM speed when all images do it at once after the Gram kernel (58 of its 81 us per image, tools/lost_finish_trace.py:
    // 6.1 TB/s); the other phases are latency.  It is launched as a programmatic dependent of the Gram kernel (which
    // triggers at its start) and CTA b waits on image b's completion counter, which the Gram epilogue releases per tile:
    // with 512-thread CTAs it cannot be co-resident with the 8-converter-warp Gram (registers), so its CTAs start as Gram
    // CTAs retire — the launch latency is hidden, the ~0.1 ms of work is not.
    // Tried (B200P_LOST_SPLIT_PCT / B200P_LOST_SMALL_THREADS keep the experiment): the first 25-100 % of the images in
    // 128-192-thread CTAs that DO fit beside a Gram CTA (4-6 warps x 62 registers next to 18 x 80) and work in the shadow
    // of the Gram of the following images, the rest in full-size CTAs afterwards.  Co-residency works (first finish CTA
    // ready 22 us after the Gram starts), but a small CTA next to a Gram CTA that saturates the shared-memory pipe needs
    // ~300 us per image and the Gram kernel loses 15-25 us: 0.501 ms at best (50 %, 192 threads) against 0.503 ms for the
    // single launch, 0.56-0.73 ms for 40-100 % at 128 threads.
    int rc = lost_gram_run(gp, 0, gp.n_tiles2, nullptr, d_degree, st); if (rc) return rc;
    static const int fin_env = [] { const char* e = getenv("B200P_LOST_FINISH"); return e ? atoi(e) : 1; }();      // experiments: 0 skip, 2 one launch after the Gram kernel
    static const int split_pct = [] { const char* e = getenv("B200P_LOST_SPLIT_PCT"); return e ? atoi(e) : 100; }();
    static const int small_threads = [] { const char* e = getenv("B200P_LOST_SMALL_THREADS"); const int v = e ? atoi(e) : kFinThreads; return v >= 64 && v <= kFinThreads ? (v / 32) * 32 : kFinThreads; }();
    if (fin_env == 0) return B200P_OK;
    int split = fin_env == 1 ? (int)((long long)n_images * split_pct / 100) : 0;                 // images of the programmatic dependent launch
    if (split < 0) split = 0; if (split > n_images) split = n_images;
    g_last_done = gp.d_done;
    if (split > 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)split); cfg.blockDim = dim3((unsigned)small_threads); cfg.dynamicSmemBytes = fin_smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        B200P_CUDA(cudaLaunchKernelEx(&cfg, k_lost_finish<true>, (const LostImageDev*)d_meta, (const float*)nullptr, (const int*)d_degree, k_patches, n_max,
                                      d_seed, d_box, d_status, (float*)nullptr, d_feats, (long long)row_stride, d, vec ? 1 : 0,
                                      (const unsigned int*)gp.d_done, 0));
    }
    if (split < n_images) {
        k_lost_finish<true><<<n_images - split, kFinThreads, fin_smem, st>>>(d_meta, nullptr, d_degree, k_patches, n_max, d_seed, d_box, d_status, nullptr,
                                                                             d_feats, (long long)row_stride, d, vec ? 1 : 0, gp.d_done, split);
        B200P_LAUNCH_CHECK("k_lost_finish");
    }
    return B200P_OK;
}

extern "C" int b200p_lost_patch_scoring(int device, const float* d_A, int n, int64_t lda, float threshold,
                                        int32_t* d_degree, int64_t* d_sel, void* stream) {
    B200P_REQUIRE(d_A && d_degree && d_sel && n >= 1 && lda >= n, B200P_EINVAL, "lost_patch_scoring: bad argument");
    B200P_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    k_lost_degree<<<(n + 7) / 8, 256, 0, st>>>(d_A, n, (long long)lda, threshold, d_degree);
    B200P_LAUNCH_CHECK("k_lost_degree");
    k_lost_rank_by_degree<<<(n + 255) / 256, 256, 0, st>>>(d_degree, n, (long long*)d_sel);
    B200P_LAUNCH_CHECK("k_lost_rank_by_degree");
    return B200P_OK;
}

extern "C" int b200p_lost_detect_box(int device, const float* d_M, int dim0, int dim1, int seed, float scale0, float scale1,
                                     int img_h, int img_w, float* d_box, int32_t* d_feat_box, int32_t* d_status, void* stream) {
    B200P_REQUIRE(d_M && d_box && d_feat_box && d_status, B200P_EINVAL, "lost_detect_box: null argument");
    const long long n = (long long)dim0 * dim1;
    B200P_REQUIRE(dim0 >= 1 && dim1 >= 1 && n <= kLostMaxPatches, B200P_EINVAL, "lost_detect_box: dims must give 1..4096 patches");
    B200P_REQUIRE(seed >= 0 && seed < n, B200P_EINVAL, "lost_detect_box: seed out of range");
    B200P_CUDA(cudaSetDevice(device));
    k_lost_detect_box<<<1, kFinThreads, 0, (cudaStream_t)stream>>>(d_M, (int)n, dim0, dim1, seed, scale0, scale1, img_h, img_w,
                                                                  d_box, d_feat_box, d_status);
    B200P_LAUNCH_CHECK("k_lost_detect_box");
    return B200P_OK;
}
