// score.cu — K1: SNIP importance score accumulation.
//   SCORE (=|+=) |W * G|      reference: train.py:258-261 (|g|), train.py:289 (|w| * |g|)
// |w|*|g| == |w*g| bit-for-bit in IEEE fp32 (one correctly rounded multiply, sign-symmetric),
// so one multiply and one sign clear replace the reference's abs, clone, abs, mul.
// HBM-bound stream: 16 B / parameter / mini-batch (read w, g, score; write score).
#include "common.cuh"

namespace b200p {

template <bool ACCUMULATE, bool VEC>
__global__ void __launch_bounds__(kThreads)
k_score_accumulate(const int32_t* __restrict__ chunk_n, ChunkTab w_tab, ChunkTab g_tab, ChunkTab s_tab,
                   int64_t c_begin, int64_t c_end) {
    const int tid = threadIdx.x;
    int64_t c = c_begin + blockIdx.x;
    if (c >= c_end) return;
    const float* w = chunk_ptr<const float>(w_tab, c);
    const float* g = chunk_ptr<const float>(g_tab, c);
    float* s = chunk_ptr<float>(s_tab, c);
    int n = __ldg(chunk_n + c);
    while (true) {
        // next chunk's addresses are fetched while this chunk streams
        const int64_t cn = c + gridDim.x;
        const bool more = cn < c_end;
        const float* wn = nullptr; const float* gn = nullptr; float* sn = nullptr; int nn = 0;
        if (more) {
            wn = chunk_ptr<const float>(w_tab, cn); gn = chunk_ptr<const float>(g_tab, cn);
            sn = chunk_ptr<float>(s_tab, cn); nn = __ldg(chunk_n + cn);
        }
        if (VEC && n == kChunk) {
            float4 wv[kVecPerThread], gv[kVecPerThread], sv4[kVecPerThread];
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                const int e = 4 * (j * kThreads + tid);
                wv[j] = ld_nc_f4(w + e);
                gv[j] = ld_nc_f4(g + e);
                if (ACCUMULATE) sv4[j] = ld_f4(s + e);
            }
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                const int e = 4 * (j * kThreads + tid);
                float4 r;
                r.x = fabsf(__fmul_rn(wv[j].x, gv[j].x));
                r.y = fabsf(__fmul_rn(wv[j].y, gv[j].y));
                r.z = fabsf(__fmul_rn(wv[j].z, gv[j].z));
                r.w = fabsf(__fmul_rn(wv[j].w, gv[j].w));
                if (ACCUMULATE) {
                    r.x = __fadd_rn(sv4[j].x, r.x); r.y = __fadd_rn(sv4[j].y, r.y);
                    r.z = __fadd_rn(sv4[j].z, r.z); r.w = __fadd_rn(sv4[j].w, r.w);
                }
                st_f4(s + e, r);
            }
        } else {
            for (int e = tid; e < n; e += kThreads) {
                float r = fabsf(__fmul_rn(w[e], g[e]));
                if (ACCUMULATE) r = __fadd_rn(s[e], r);
                s[e] = r;
            }
        }
        if (!more) break;
        c = cn; w = wn; g = gn; s = sn; n = nn;
    }
}

// Multi-batch form: SCORE (=|+=) |W*G_0| + |W*G_1| + ... + |W*G_{B-1}|, added left to right, which is
// bit-identical to B successive k_score_accumulate launches but reads W once and touches SCORE once:
// 4*(B+2) B/param (+4 when accumulating) instead of 16*B.  The B200's 180 GB make it free to keep the
// gradient sets of all SNIP mini-batches resident (102 MB each for ResNet-50) and fold them in one pass.
template <bool ACCUMULATE, int B>
__global__ void __launch_bounds__(kThreads)
k_score_multi(const int32_t* __restrict__ chunk_n, ChunkTab w_tab, GradTabs g_tabs, ChunkTab s_tab,
              int64_t c_begin, int64_t c_end, int vec_ok) {
    const int tid = threadIdx.x;
    for (int64_t c = c_begin + blockIdx.x; c < c_end; c += gridDim.x) {
        const int n = __ldg(chunk_n + c);
        const float* __restrict__ w = chunk_ptr<const float>(w_tab, c);
        float* __restrict__ s = chunk_ptr<float>(s_tab, c);
        const float* g[B];
#pragma unroll
        for (int b = 0; b < B; ++b) g[b] = chunk_ptr<const float>(g_tabs.t[b], c);
        if (vec_ok && n == kChunk) {
#pragma unroll 2
            for (int j = 0; j < kVecPerThread; ++j) {
                const int e = 4 * (j * kThreads + tid);
                const float4 wv = ld_nc_f4(w + e);
                float4 gv[B];
#pragma unroll
                for (int b = 0; b < B; ++b) gv[b] = ld_nc_f4(g[b] + e);
                float4 r;
                if (ACCUMULATE) r = ld_f4(s + e);
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const float4 t = make_float4(fabsf(__fmul_rn(wv.x, gv[b].x)), fabsf(__fmul_rn(wv.y, gv[b].y)),
                                                 fabsf(__fmul_rn(wv.z, gv[b].z)), fabsf(__fmul_rn(wv.w, gv[b].w)));
                    if (b == 0 && !ACCUMULATE) r = t;
                    else { r.x = __fadd_rn(r.x, t.x); r.y = __fadd_rn(r.y, t.y); r.z = __fadd_rn(r.z, t.z); r.w = __fadd_rn(r.w, t.w); }
                }
                st_f4(s + e, r);
            }
        } else {
            for (int e = tid; e < n; e += kThreads) {
                float r = ACCUMULATE ? s[e] : 0.f;
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const float t = fabsf(__fmul_rn(w[e], g[b][e]));
                    r = (b == 0 && !ACCUMULATE) ? t : __fadd_rn(r, t);
                }
                s[e] = r;
            }
        }
    }
}

// dst[i] = ((src[0][i] + src[1][i]) + src[2][i]) + ...   part p at src + p * part_stride.
// Fixed left-to-right order: the multi-GPU score exchange sums the ranks' partial scores in rank
// (= mini-batch) order on every rank count, so the result does not depend on the collective's
// reduction topology (SURVEY §7 "cross-GPU-count determinism").
__global__ void __launch_bounds__(kThreads)
k_sum_parts(float* __restrict__ dst, const float* __restrict__ src, int n_parts, int64_t part_stride,
            int64_t n, int vec_ok) {
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * kThreads;
    int64_t done = 0;
    if (vec_ok) {
        const int64_t nvec = n >> 2;
        for (int64_t v = tid; v < nvec; v += nthreads) {
            float4 acc = ld_nc_f4(src + 4 * v);
            for (int p = 1; p < n_parts; ++p) {
                const float4 x = ld_nc_f4(src + p * part_stride + 4 * v);
                acc.x = __fadd_rn(acc.x, x.x); acc.y = __fadd_rn(acc.y, x.y);
                acc.z = __fadd_rn(acc.z, x.z); acc.w = __fadd_rn(acc.w, x.w);
            }
            st_f4(dst + 4 * v, acc);
        }
        done = nvec << 2;
    }
    for (int64_t i = done + tid; i < n; i += nthreads) {
        float acc = src[i];
        for (int p = 1; p < n_parts; ++p) acc = __fadd_rn(acc, src[p * part_stride + i]);
        dst[i] = acc;
    }
}

}  // namespace b200p

using namespace b200p;

template <bool ACC>
static void launch_multi(int nb, int grid, cudaStream_t st, const int32_t* cn, ChunkTab w, const GradTabs& g, ChunkTab s,
                         int64_t c0, int64_t c1, int vec) {
    switch (nb) {
        case 1: k_score_multi<ACC, 1><<<grid, kThreads, 0, st>>>(cn, w, g, s, c0, c1, vec); break;
        case 2: k_score_multi<ACC, 2><<<grid, kThreads, 0, st>>>(cn, w, g, s, c0, c1, vec); break;
        case 3: k_score_multi<ACC, 3><<<grid, kThreads, 0, st>>>(cn, w, g, s, c0, c1, vec); break;
        case 4: k_score_multi<ACC, 4><<<grid, kThreads, 0, st>>>(cn, w, g, s, c0, c1, vec); break;
        case 5: k_score_multi<ACC, 5><<<grid, kThreads, 0, st>>>(cn, w, g, s, c0, c1, vec); break;
        case 6: k_score_multi<ACC, 6><<<grid, kThreads, 0, st>>>(cn, w, g, s, c0, c1, vec); break;
        case 7: k_score_multi<ACC, 7><<<grid, kThreads, 0, st>>>(cn, w, g, s, c0, c1, vec); break;
        default: k_score_multi<ACC, 8><<<grid, kThreads, 0, st>>>(cn, w, g, s, c0, c1, vec); break;
    }
}

extern "C" int b200p_score_accumulate_multi(b200p_plan* p, const b200p_ptrtable* const* g_tables, int n_sets, int accumulate,
                                            int64_t chunk_begin, int64_t chunk_end, void* stream) {
    B200P_REQUIRE(p != nullptr && g_tables != nullptr, B200P_EINVAL, "score_accumulate_multi: null argument");
    B200P_REQUIRE(n_sets >= 1, B200P_EINVAL, "score_accumulate_multi: need at least one gradient set");
    B200P_REQUIRE(p->bound[B200P_SLOT_W] && p->bound[B200P_SLOT_SCORE], B200P_ESTATE, "score_accumulate_multi: W and SCORE slots must be bound");
    if (chunk_end < 0) chunk_end = p->n_chunks;
    B200P_REQUIRE(chunk_begin >= 0 && chunk_begin <= chunk_end && chunk_end <= p->n_chunks, B200P_EINVAL, "score_accumulate_multi: bad chunk range");
    bool vec = p->vec_ok[B200P_SLOT_W] && p->vec_ok[B200P_SLOT_SCORE];
    for (int i = 0; i < n_sets; ++i) {
        B200P_REQUIRE(g_tables[i] != nullptr && g_tables[i]->plan == p, B200P_EINVAL, "score_accumulate_multi: table belongs to another plan");
        vec = vec && g_tables[i]->vec_ok;
    }
    if (chunk_begin == chunk_end) return B200P_OK;
    B200P_CUDA(cudaSetDevice(p->device));
    const int grid = p->grid_for(chunk_end - chunk_begin, 4);
    cudaStream_t st = (cudaStream_t)stream;
    for (int i0 = 0; i0 < n_sets; i0 += kMaxSets) {          // groups of up to 8 sets per launch
        const int nb = n_sets - i0 < kMaxSets ? n_sets - i0 : kMaxSets;
        GradTabs g;
        for (int b = 0; b < kMaxSets; ++b) g.t[b] = (ChunkTab)g_tables[i0 + (b < nb ? b : 0)]->d_tab;
        if (accumulate || i0 > 0) launch_multi<true>(nb, grid, st, p->d_chunk_n, p->tab(B200P_SLOT_W), g, p->tab(B200P_SLOT_SCORE), chunk_begin, chunk_end, vec ? 1 : 0);
        else                      launch_multi<false>(nb, grid, st, p->d_chunk_n, p->tab(B200P_SLOT_W), g, p->tab(B200P_SLOT_SCORE), chunk_begin, chunk_end, vec ? 1 : 0);
        B200P_LAUNCH_CHECK("k_score_multi");
    }
    return B200P_OK;
}

extern "C" int b200p_sum_parts(int device, float* d_dst, const float* d_src, int n_parts, int64_t part_stride,
                               int64_t n, void* stream) {
    B200P_REQUIRE(d_dst != nullptr && d_src != nullptr, B200P_EINVAL, "sum_parts: null argument");
    B200P_REQUIRE(n_parts >= 1 && n >= 0 && part_stride >= n, B200P_EINVAL, "sum_parts: bad sizes");
    if (n == 0) return B200P_OK;
    B200P_CUDA(cudaSetDevice(device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const bool vec = (((uintptr_t)d_dst | (uintptr_t)d_src) & 15u) == 0 && (part_stride & 3) == 0;
    int64_t blocks = (n / 4 + kThreads - 1) / kThreads;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    if (blocks < 1) blocks = 1;
    k_sum_parts<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(d_dst, d_src, n_parts, part_stride, n, vec ? 1 : 0);
    B200P_LAUNCH_CHECK("k_sum_parts");
    return B200P_OK;
}

extern "C" int b200p_score_accumulate(b200p_plan* p, int accumulate, int64_t chunk_begin,
                                      int64_t chunk_end, void* stream) {
    B200P_REQUIRE(p != nullptr, B200P_EINVAL, "score_accumulate: null plan");
    B200P_REQUIRE(p->bound[B200P_SLOT_W] && p->bound[B200P_SLOT_G] && p->bound[B200P_SLOT_SCORE],
                  B200P_ESTATE, "score_accumulate: W, G and SCORE slots must be bound");
    if (chunk_end < 0) chunk_end = p->n_chunks;
    B200P_REQUIRE(chunk_begin >= 0 && chunk_begin <= chunk_end && chunk_end <= p->n_chunks,
                  B200P_EINVAL, "score_accumulate: bad chunk range");
    if (chunk_begin == chunk_end) return B200P_OK;
    B200P_CUDA(cudaSetDevice(p->device));
    const bool vec = p->vec_ok[B200P_SLOT_W] && p->vec_ok[B200P_SLOT_G] && p->vec_ok[B200P_SLOT_SCORE];
    const int grid = p->grid_for(chunk_end - chunk_begin, 4);
    cudaStream_t st = (cudaStream_t)stream;
    ChunkTab w = p->tab(B200P_SLOT_W), g = p->tab(B200P_SLOT_G), s = p->tab(B200P_SLOT_SCORE);
    const int32_t* cn = p->d_chunk_n;
    if (accumulate) {
        if (vec) k_score_accumulate<true, true><<<grid, kThreads, 0, st>>>(cn, w, g, s, chunk_begin, chunk_end);
        else     k_score_accumulate<true, false><<<grid, kThreads, 0, st>>>(cn, w, g, s, chunk_begin, chunk_end);
    } else {
        if (vec) k_score_accumulate<false, true><<<grid, kThreads, 0, st>>>(cn, w, g, s, chunk_begin, chunk_end);
        else     k_score_accumulate<false, false><<<grid, kThreads, 0, st>>>(cn, w, g, s, chunk_begin, chunk_end);
    }
    B200P_LAUNCH_CHECK("k_score_accumulate");
    return B200P_OK;
}
