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
};

__device__ __forceinline__ void sgd_elem(float& w, float g, float& buf, bool on, const SgdArgs& a, float& weff) {
    g = on ? g : 0.f;                                       // MulBackward of the mask
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
                sgd_elem(wv[j].x, gv[j].x, bv[j].x, nib[j] & 1u, a, o.x);
                sgd_elem(wv[j].y, gv[j].y, bv[j].y, nib[j] & 2u, a, o.y);
                sgd_elem(wv[j].z, gv[j].z, bv[j].z, nib[j] & 4u, a, o.z);
                sgd_elem(wv[j].w, gv[j].w, bv[j].w, nib[j] & 8u, a, o.w);
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
                sgd_elem(wv, g[e], bv, on, a, o);
                w[e] = wv;
                if (use_buf) buf[e] = bv;
                if (we) we[e] = o;
                if (wh) wh[e] = __float2bfloat16_rn(o);
            }
        }
    }
}

}  // namespace b200p

using namespace b200p;

extern "C" int b200p_masked_sgd_step(b200p_plan* p, const uint32_t* d_mask, float lr, float momentum,
                                     float dampening, float weight_decay, int flags, void* stream) {
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
    bool vec = p->vec_ok[B200P_SLOT_W] && p->vec_ok[B200P_SLOT_G];
    if (momentum != 0.f) vec = vec && p->vec_ok[B200P_SLOT_BUF];
    if (flags & B200P_SGD_EMIT_WEFF) vec = vec && p->vec_ok[B200P_SLOT_WEFF];
    if (flags & B200P_SGD_EMIT_WEFF16) vec = vec && p->vec_ok[B200P_SLOT_WEFF16];
    a.vec_ok = vec ? 1 : 0;
    k_masked_sgd<<<p->grid_for(p->n_chunks, 3), kThreads, 0, (cudaStream_t)stream>>>(a, p->n_chunks);
    B200P_LAUNCH_CHECK("k_masked_sgd");
    return B200P_OK;
}
