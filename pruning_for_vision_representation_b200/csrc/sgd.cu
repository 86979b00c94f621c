// sgd.cu — K4: fused masked-weight apply + masked SGD/momentum update.
//
// One pass replaces, per prunable tensor and step, the reference's
//   forward pre-hook   weight = weight_mask * weight_orig        torch/nn/utils/prune.py:71-74
//   MulBackward        grad(weight_orig) = grad(weight) * mask   (autograd of the line above)
//   torch.optim.SGD    g += wd*p; buf = mu*buf + (1-damp)*g; g = nesterov ? g + mu*buf : buf;
//                      p -= lr*g                                  torch/optim/sgd.py:343-380
// and emits the next step's masked weight (fp32 and/or bf16) so pruned weights and their
// gradients never re-densify.  As in the reference, pruned entries of weight_orig keep
// decaying (weight decay) and keep a momentum buffer; they just never reach the forward.
// 22.125 B / parameter / step with the bf16 emit (read w,g,buf,bit; write w,buf,bf16).
#include "common.cuh"

namespace b200p {

struct SgdArgs {
    const int32_t* chunk_n;
    ChunkTab w_tab, g_tab, buf_tab, weff_tab, weff16_tab;
    const uint32_t* mask;      // nullable: dense SGD
    float lr, momentum, one_minus_damp, wd;
    int flags, vec_ok;
    // device-side control block (nullable): ctl[0] multiplies every gradient before use — loss-scale removal
    // (GradScaler.unscale_), global-norm clipping coefficient (clip_grad_norm_) and the 1/world of a gradient
    // all-reduce folded into one factor; ctl[1] != 0 skips the whole step (a non-finite gradient was found,
    // GradScaler.step).  Decided on the device: no host round trip between backward and step.  train.py:54-66
    const float* ctl;
};

__device__ __forceinline__ void sgd_elem(float& w, float g, float& buf, bool on, const SgdArgs& a, float& weff, float gmul) {
    g = on ? g * gmul : 0.f;                                // MulBackward of the mask (+ unscale / clip factor)
    if (a.wd != 0.f) g = fmaf(a.wd, w, g);                  // grad.add(param, alpha=wd)
    if (a.momentum != 0.f) {
        if (a.flags & B200P_SGD_FIRST_STEP) buf = g;        // buf = clone(grad)
        else {
            const float t = __fmul_rn(buf, a.momentum);     // buf.mul_(momentum)
            buf = a.one_minus_damp == 1.f ? __fadd_rn(t, g) : fmaf(a.one_minus_damp, g, t);
        }
        g = (a.flags & B200P_SGD_NESTEROV) ? fmaf(a.momentum, buf, g) : buf;
    }
    w = fmaf(-a.lr, g, w);                                  // param.add_(grad, alpha=-lr)
    weff = on ? w : 0.f;
}

__global__ void __launch_bounds__(kThreads)
k_masked_sgd(SgdArgs a, int64_t n_chunks) {
    const int tid = threadIdx.x;
    const bool use_buf = a.momentum != 0.f;
    const bool read_buf = use_buf && !(a.flags & B200P_SGD_FIRST_STEP);
    const bool want32 = a.flags & B200P_SGD_EMIT_WEFF, want16 = a.flags & B200P_SGD_EMIT_WEFF16;
    float gmul = 1.f;
    if (a.ctl) {
        if (__ldg(a.ctl + 1) != 0.f) return;               // skipped step: weights, momentum and the emitted weights stay as they are
        gmul = __ldg(a.ctl);
    }
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int n = __ldg(a.chunk_n + c);
        float* __restrict__ w = chunk_ptr<float>(a.w_tab, c);
        const float* __restrict__ g = chunk_ptr<const float>(a.g_tab, c);
        float* __restrict__ buf = use_buf ? chunk_ptr<float>(a.buf_tab, c) : nullptr;
        float* __restrict__ we = want32 ? chunk_ptr<float>(a.weff_tab, c) : nullptr;
        __nv_bfloat16* __restrict__ wh = want16 ? chunk_ptr<__nv_bfloat16>(a.weff16_tab, c) : nullptr;
        const uint32_t* m = a.mask ? a.mask + c * kWordsPerChunk : nullptr;
        if (a.vec_ok && n == kChunk) {
            float4 wv[kVecPerThread], gv[kVecPerThread], bv[kVecPerThread];
            uint32_t nib[kVecPerThread];
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                const int e = 4 * (j * kThreads + tid);
                wv[j] = ld_f4(w + e);
                gv[j] = ld_nc_f4(g + e);
                bv[j] = read_buf ? ld_f4(buf + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                nib[j] = m ? nibble_of(__ldg(m + vec_word_index(j))) : 0xFu;
            }
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                const int e = 4 * (j * kThreads + tid);
                float4 o;
                sgd_elem(wv[j].x, gv[j].x, bv[j].x, nib[j] & 1u, a, o.x, gmul);
                sgd_elem(wv[j].y, gv[j].y, bv[j].y, nib[j] & 2u, a, o.y, gmul);
                sgd_elem(wv[j].z, gv[j].z, bv[j].z, nib[j] & 4u, a, o.z, gmul);
                sgd_elem(wv[j].w, gv[j].w, bv[j].w, nib[j] & 8u, a, o.w, gmul);
                st_f4(w + e, wv[j]);
                if (use_buf) st_f4(buf + e, bv[j]);
                if (we) st_f4(we + e, o);
                if (wh) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                    uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(wh + e) = pk;
                }
            }
        } else {
            for (int e = tid; e < n; e += kThreads) {
                const bool on = m ? ((__ldg(m + (e >> 5)) >> (e & 31)) & 1u) : true;
                float wv = w[e], bv = read_buf ? buf[e] : 0.f, o;
                sgd_elem(wv, g[e], bv, on, a, o, gmul);
                w[e] = wv;
                if (use_buf) buf[e] = bv;
                if (we) we[e] = o;
                if (wh) wh[e] = __float2bfloat16_rn(o);
            }
        }
    }
}

// ---- gradient statistics for clipping / loss scaling -------------------------------------------------------------------
// out[0] += sum over kept entries of g^2 (what weight_orig.grad holds after the reference's MulBackward, train.py:57-66
// clip_grad_norm_), out[1] += number of non-finite gradient entries, kept or not (inf * 0 = NaN in the reference's masked
// gradient, so GradScaler's inf check sees those too).  One read of the gradients, 4.125 B/param.
__global__ void __launch_bounds__(kThreads)
k_grad_stats(const int32_t* __restrict__ chunk_n, ChunkTab g_tab, const uint32_t* __restrict__ mask, double* __restrict__ out,
             int64_t n_chunks, int vec_ok) {
    __shared__ double s_sq, s_bad;
    const int tid = threadIdx.x;
    if (tid == 0) { s_sq = 0.0; s_bad = 0.0; }
    __syncthreads();
    double sq = 0.0; unsigned bad = 0;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int n = __ldg(chunk_n + c);
        const float* __restrict__ g = chunk_ptr<const float>(g_tab, c);
        const uint32_t* m = mask ? mask + c * kWordsPerChunk : nullptr;
        float part = 0.f;
        if (vec_ok && n == kChunk) {
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                const float4 v = ld_nc_f4(g + 4 * (j * kThreads + tid));
                const uint32_t nib = m ? nibble_of(__ldg(m + vec_word_index(j))) : 0xFu;
                const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    bad += (__float_as_uint(f[q]) & 0x7F800000u) == 0x7F800000u ? 1u : 0u;
                    if ((nib >> q) & 1u) part = fmaf(f[q], f[q], part);
                }
            }
        } else {
            for (int e = tid; e < n; e += kThreads) {
                const float v = g[e];
                const bool on = m ? ((__ldg(m + (e >> 5)) >> (e & 31)) & 1u) : true;
                bad += (__float_as_uint(v) & 0x7F800000u) == 0x7F800000u ? 1u : 0u;
                if (on) part = fmaf(v, v, part);
            }
        }
        sq += (double)part;                                 // 16 squares per thread in fp32, everything above in fp64
    }
    double b = (double)bad;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); }
    if ((tid & 31) == 0) { atomicAdd(&s_sq, sq); atomicAdd(&s_bad, b); }
    __syncthreads();
    if (tid == 0) { atomicAdd(out + 0, s_sq); if (s_bad != 0.0) atomicAdd(out + 1, s_bad); }
}

// ---- EMA of the master weights (utils.py:159-170 ExponentialMovingAverage over weight_orig) ---------------------------
// ema = copy ? w : decay * ema + (1 - decay) * w, the arithmetic of the reference's ema_avg; 12 B/param (8 for a copy).
__global__ void __launch_bounds__(kThreads)
k_ema_update(const int32_t* __restrict__ chunk_n, ChunkTab w_tab, ChunkTab ema_tab, float decay, int copy, int64_t n_chunks, int vec_ok) {
    const int tid = threadIdx.x;
    const float omd = 1.f - decay;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int n = __ldg(chunk_n + c);
        const float* __restrict__ w = chunk_ptr<const float>(w_tab, c);
        float* __restrict__ e = chunk_ptr<float>(ema_tab, c);
        if (vec_ok && n == kChunk) {
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                const int i = 4 * (j * kThreads + tid);
                const float4 wv = ld_nc_f4(w + i);
                float4 o = wv;
                if (!copy) {
                    const float4 ev = ld_f4(e + i);
                    o.x = __fadd_rn(__fmul_rn(decay, ev.x), __fmul_rn(omd, wv.x)); o.y = __fadd_rn(__fmul_rn(decay, ev.y), __fmul_rn(omd, wv.y));
                    o.z = __fadd_rn(__fmul_rn(decay, ev.z), __fmul_rn(omd, wv.z)); o.w = __fadd_rn(__fmul_rn(decay, ev.w), __fmul_rn(omd, wv.w));
                }
                st_f4(e + i, o);
            }
        } else {
            for (int i = tid; i < n; i += kThreads)
                e[i] = copy ? w[i] : __fadd_rn(__fmul_rn(decay, e[i]), __fmul_rn(omd, w[i]));
        }
    }
}

}  // namespace b200p

using namespace b200p;

extern "C" int b200p_grad_stats(b200p_plan* p, const uint32_t* d_mask, double* d_out2, void* stream) {
    B200P_REQUIRE(p != nullptr && d_out2 != nullptr, B200P_EINVAL, "grad_stats: null argument");
    B200P_REQUIRE(p->bound[B200P_SLOT_G], B200P_ESTATE, "grad_stats: G slot must be bound");
    B200P_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    B200P_CUDA(cudaMemsetAsync(d_out2, 0, 2 * sizeof(double), st));
    k_grad_stats<<<p->grid_for(p->n_chunks, 4), kThreads, 0, st>>>(p->d_chunk_n, p->tab(B200P_SLOT_G), d_mask, d_out2, p->n_chunks,
                                                                  p->vec_ok[B200P_SLOT_G] ? 1 : 0);
    B200P_LAUNCH_CHECK("k_grad_stats");
    return B200P_OK;
}

extern "C" int b200p_ema_update(b200p_plan* p, float decay, int copy, void* stream) {
    B200P_REQUIRE(p != nullptr, B200P_EINVAL, "ema_update: null plan");
    B200P_REQUIRE(p->bound[B200P_SLOT_W] && p->bound[B200P_SLOT_EMA], B200P_ESTATE, "ema_update: W and EMA slots must be bound");
    B200P_CUDA(cudaSetDevice(p->device));
    k_ema_update<<<p->grid_for(p->n_chunks, 4), kThreads, 0, (cudaStream_t)stream>>>(p->d_chunk_n, p->tab(B200P_SLOT_W), p->tab(B200P_SLOT_EMA), decay,
                                                                                     copy ? 1 : 0, p->n_chunks,
                                                                                     (p->vec_ok[B200P_SLOT_W] && p->vec_ok[B200P_SLOT_EMA]) ? 1 : 0);
    B200P_LAUNCH_CHECK("k_ema_update");
    return B200P_OK;
}

extern "C" int b200p_masked_sgd_step(b200p_plan* p, const uint32_t* d_mask, float lr, float momentum,
                                     float dampening, float weight_decay, int flags, void* stream) {
    return b200p_masked_sgd_step_ctl(p, d_mask, lr, momentum, dampening, weight_decay, flags, nullptr, stream);
}

extern "C" int b200p_masked_sgd_step_ctl(b200p_plan* p, const uint32_t* d_mask, float lr, float momentum,
                                         float dampening, float weight_decay, int flags, const float* d_ctl, void* stream) {
    B200P_REQUIRE(p != nullptr, B200P_EINVAL, "masked_sgd_step: null plan");
    B200P_REQUIRE(p->bound[B200P_SLOT_W] && p->bound[B200P_SLOT_G], B200P_ESTATE, "masked_sgd_step: W and G slots must be bound");
    if (momentum != 0.f) B200P_REQUIRE(p->bound[B200P_SLOT_BUF], B200P_ESTATE, "masked_sgd_step: BUF slot must be bound when momentum != 0");
    if (flags & B200P_SGD_EMIT_WEFF) B200P_REQUIRE(p->bound[B200P_SLOT_WEFF], B200P_ESTATE, "masked_sgd_step: WEFF slot is not bound");
    if (flags & B200P_SGD_EMIT_WEFF16) B200P_REQUIRE(p->bound[B200P_SLOT_WEFF16], B200P_ESTATE, "masked_sgd_step: WEFF16 slot is not bound");
    if (flags & B200P_SGD_NESTEROV) B200P_REQUIRE(momentum > 0.f && dampening == 0.f, B200P_EINVAL, "masked_sgd_step: nesterov needs momentum > 0 and zero dampening");
    B200P_CUDA(cudaSetDevice(p->device));
    SgdArgs a;
    a.chunk_n = p->d_chunk_n;
    a.w_tab = p->tab(B200P_SLOT_W); a.g_tab = p->tab(B200P_SLOT_G);
    a.buf_tab = p->tab(B200P_SLOT_BUF); a.weff_tab = p->tab(B200P_SLOT_WEFF);
    a.weff16_tab = p->tab(B200P_SLOT_WEFF16);
    a.mask = d_mask; a.lr = lr; a.momentum = momentum; a.one_minus_damp = 1.f - dampening; a.wd = weight_decay; a.flags = flags;
    a.ctl = d_ctl;
    bool vec = p->vec_ok[B200P_SLOT_W] && p->vec_ok[B200P_SLOT_G];
    if (momentum != 0.f) vec = vec && p->vec_ok[B200P_SLOT_BUF];
    if (flags & B200P_SGD_EMIT_WEFF) vec = vec && p->vec_ok[B200P_SLOT_WEFF];
    if (flags & B200P_SGD_EMIT_WEFF16) vec = vec && p->vec_ok[B200P_SLOT_WEFF16];
    a.vec_ok = vec ? 1 : 0;
    k_masked_sgd<<<p->grid_for(p->n_chunks, 3), kThreads, 0, (cudaStream_t)stream>>>(a, p->n_chunks);
    B200P_LAUNCH_CHECK("k_masked_sgd");
    return B200P_OK;
}
