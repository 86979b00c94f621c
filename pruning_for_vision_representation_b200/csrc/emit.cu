// emit.cu — K3 mask emit, K5 sparsity count, mask format conversions, mask apply.
//
// K3 replaces train.py:311-317 (mask = (score > thr).float(); custom_from_mask) and the
// index_put of torch/nn/utils/prune.py:538 + per-tensor slicing :1149-1161.  The kernel-side
// truth is the bit-packed mask (1/8 B per parameter); the fp32 `weight_mask` buffers and the
// masked `weight` tensors that the reference's checkpoint format needs are optional fused
// outputs of the same pass.
#include "common.cuh"
#include "emit_body.cuh"

namespace b200p {

__global__ void __launch_bounds__(kThreads)
k_emit_masks(EmitArgs a, int64_t c_begin, int64_t c_end) {
    if (a.patch && a.st->prov_ok) {
        emit_patch_body(a.chunk_n, a.key_tab, a.old_mask, a.st, a.cand_key, a.cand_pos, a.prov, a.mode, a.n_chunks,
                        patch_vals_from_state(a.st, a.mode));
        return;
    }
    emit_full_body(a, c_begin, c_end);
}

// ---- K5: zeros of the effective weight ---------------------------------------------------
// out[0] += #(bit == 0 || w == 0)   out[1] += #(bit == 1)
__global__ void __launch_bounds__(kThreads)
k_count_zeros(const int32_t* __restrict__ chunk_n, ChunkTab w_tab, const uint32_t* __restrict__ mask,
              unsigned long long* __restrict__ out, int64_t n_chunks, int vec_ok, int use_weights) {
    __shared__ unsigned long long s_z, s_b;
    if (threadIdx.x == 0) { s_z = 0; s_b = 0; }
    __syncthreads();
    const int tid = threadIdx.x;
    unsigned long long zeros = 0, bits = 0;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int n = __ldg(chunk_n + c);
        const uint32_t* m = mask ? mask + c * kWordsPerChunk : nullptr;
        if (!use_weights) {
            // packed-only fast path: zeros = n - popcount
            if (tid < kWordsPerChunk) {
                const uint32_t w = m ? __ldg(m + tid) : 0u;
                bits += __popc(w);
            }
            if (tid == 0) zeros += n;        // corrected below by subtracting the bits
            continue;
        }
        const float* __restrict__ w = chunk_ptr<const float>(w_tab, c);
        if (vec_ok && n == kChunk) {
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                const float4 v = ld_nc_f4(w + 4 * (j * kThreads + tid));
                uint32_t nib = 0xFu;
                if (m) nib = nibble_of(__ldg(m + vec_word_index(j)));
                bits += __popc(nib);
                zeros += ((nib & 1u) == 0 || v.x == 0.f) + ((nib & 2u) == 0 || v.y == 0.f) +
                         ((nib & 4u) == 0 || v.z == 0.f) + ((nib & 8u) == 0 || v.w == 0.f);
            }
        } else {
            for (int e = tid; e < n; e += kThreads) {
                uint32_t bit = 1u;
                if (m) bit = (__ldg(m + (e >> 5)) >> (e & 31)) & 1u;
                bits += bit;
                zeros += (bit == 0 || w[e] == 0.f);
            }
        }
    }
    if (!use_weights) zeros = zeros - bits;   // per-thread partials may wrap; the 64-bit sum is exact
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        zeros += __shfl_xor_sync(0xFFFFFFFFu, zeros, o);
        bits += __shfl_xor_sync(0xFFFFFFFFu, bits, o);
    }
    if ((tid & 31) == 0) { atomicAdd(&s_z, zeros); atomicAdd(&s_b, bits); }
    __syncthreads();
    if (tid == 0) { atomicAdd(out + 0, s_z); atomicAdd(out + 1, s_b); }
}

// ---- conversions ---------------------------------------------------------------------------
// DIR 0: MASKF (fp32 0/1) -> packed.   DIR 1: packed -> MASKF.
// DIR 2: WEFF = bit ? W : 0 (and/or bf16).   DIR 3: G = bit ? G : 0.
template <int DIR>
__global__ void __launch_bounds__(kThreads)
k_mask_convert(const int32_t* __restrict__ chunk_n, ChunkTab f_tab, ChunkTab w_tab, ChunkTab h_tab,
               uint32_t* __restrict__ mask, int64_t n_chunks, int vec_ok, int outputs) {
    const int tid = threadIdx.x;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int n = __ldg(chunk_n + c);
        uint32_t* m = mask + c * kWordsPerChunk;
        float* f = f_tab ? chunk_ptr<float>(f_tab, c) : nullptr;
        const float* w = w_tab ? chunk_ptr<const float>(w_tab, c) : nullptr;
        __nv_bfloat16* h = h_tab ? chunk_ptr<__nv_bfloat16>(h_tab, c) : nullptr;
        if (vec_ok && n == kChunk) {
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                const int e = 4 * (j * kThreads + tid);
                if (DIR == 0) {
                    const float4 v = ld_nc_f4(f + e);
                    const uint32_t nib = (v.x != 0.f ? 1u : 0u) | (v.y != 0.f ? 2u : 0u) | (v.z != 0.f ? 4u : 0u) | (v.w != 0.f ? 8u : 0u);
                    const uint32_t word = gather_nibbles(nib);
                    if ((tid & 7) == 0) m[vec_word_index(j)] = word;
                } else {
                    const uint32_t nib = nibble_of(__ldg(m + vec_word_index(j)));
                    if (DIR == 1) {
                        float4 o; o.x = (nib & 1u) ? 1.f : 0.f; o.y = (nib & 2u) ? 1.f : 0.f; o.z = (nib & 4u) ? 1.f : 0.f; o.w = (nib & 8u) ? 1.f : 0.f;
                        st_f4(f + e, o);
                    } else if (DIR == 2) {
                        const float4 v = ld_nc_f4(w + e);
                        float4 o; o.x = (nib & 1u) ? v.x : 0.f; o.y = (nib & 2u) ? v.y : 0.f; o.z = (nib & 4u) ? v.z : 0.f; o.w = (nib & 8u) ? v.w : 0.f;
                        if (outputs & B200P_EMIT_WEFF) st_f4(f + e, o);
                        if (h) {
                            __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                            uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
                            *reinterpret_cast<uint2*>(h + e) = pk;
                        }
                    } else {
                        float4 v = ld_f4(f + e);
                        v.x = (nib & 1u) ? v.x : 0.f; v.y = (nib & 2u) ? v.y : 0.f; v.z = (nib & 4u) ? v.z : 0.f; v.w = (nib & 8u) ? v.w : 0.f;
                        st_f4(f + e, v);
                    }
                }
            }
        } else {
            const int warp = tid >> 5, lane = tid & 31;
            for (int wd = warp; wd < kWordsPerChunk; wd += kThreads / 32) {
                const int e = wd * 32 + lane;
                if (DIR == 0) {
                    const bool on = e < n && f[e] != 0.f;
                    const uint32_t word = __ballot_sync(0xFFFFFFFFu, on);
                    if (lane == 0) m[wd] = word;
                } else if (e < n) {
                    const bool on = (__ldg(m + wd) >> lane) & 1u;
                    if (DIR == 1) f[e] = on ? 1.f : 0.f;
                    else if (DIR == 2) {
                        const float o = on ? w[e] : 0.f;
                        if (outputs & B200P_EMIT_WEFF) f[e] = o;
                        if (h) h[e] = __float2bfloat16_rn(o);
                    } else f[e] = on ? f[e] : 0.f;
                }
            }
        }
    }
}

}  // namespace b200p

using namespace b200p;

extern "C" int b200p_emit_masks(b200p_plan* p, int key_source, int mode, int force, float forced_threshold,
                                const uint32_t* d_old_mask, uint32_t* d_new_mask, int outputs,
                                int64_t chunk_begin, int64_t chunk_end, void* stream) {
    B200P_REQUIRE(p != nullptr && d_new_mask != nullptr, B200P_EINVAL, "emit_masks: null argument");
    B200P_REQUIRE(key_source == B200P_KEY_ABS_W || key_source == B200P_KEY_SCORE, B200P_EINVAL, "emit_masks: bad key_source");
    B200P_REQUIRE(mode == B200P_MODE_SNIP_STRICT || mode == B200P_MODE_EXACT_K, B200P_EINVAL, "emit_masks: bad mode");
    B200P_REQUIRE(force >= 0 && force <= 3, B200P_EINVAL, "emit_masks: bad force");
    const int kslot = key_source == B200P_KEY_ABS_W ? B200P_SLOT_W : B200P_SLOT_SCORE;
    B200P_REQUIRE(p->bound[kslot], B200P_ESTATE, "emit_masks: key slot is not bound");
    if (outputs & B200P_EMIT_MASKF) B200P_REQUIRE(p->bound[B200P_SLOT_MASKF], B200P_ESTATE, "emit_masks: MASKF slot is not bound");
    if (outputs & B200P_EMIT_WEFF) B200P_REQUIRE(p->bound[B200P_SLOT_WEFF] && p->bound[B200P_SLOT_W], B200P_ESTATE, "emit_masks: WEFF/W slots are not bound");
    if (chunk_end < 0) chunk_end = p->n_chunks;
    B200P_REQUIRE(chunk_begin >= 0 && chunk_begin <= chunk_end && chunk_end <= p->n_chunks, B200P_EINVAL, "emit_masks: bad chunk range");
    B200P_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    B200P_CUDA(cudaMemsetAsync(&p->d_state->n_kept, 0, sizeof(unsigned long long), st));
    if (chunk_begin == chunk_end) return B200P_OK;
    // emit-by-patch: directly after b200p_select_kth with the same key source / mode / old mask, over the whole
    // parameter set and without fused fp32 outputs.  Whether the provisional mask is valid is only known on the
    // device (the select may have fallen back to histogram mode), so the full pass is still launched and exits at
    // once when the patch has done the job.
    const bool patch = p->prov_armed && force == 0 && outputs == 0 && chunk_begin == p->prov_c0 && chunk_end == p->prov_c1 &&
                       p->prov_key_source == key_source && p->prov_mode == mode && p->prov_old_mask == d_old_mask;
    uint32_t* prov = p->prov_target ? p->prov_target : p->d_prov;
    p->prov_armed = false; p->prov_target = nullptr;
    EmitArgs a;
    a.patch = patch ? 1 : 0; a.prov = prov; a.cand_key = p->d_cand_key; a.cand_pos = p->d_cand_pos; a.n_chunks = p->n_chunks;
    a.chunk_n = p->d_chunk_n;
    a.key_tab = p->tab(kslot);
    a.w_tab = p->tab(B200P_SLOT_W);
    a.maskf_tab = p->tab(B200P_SLOT_MASKF);
    a.weff_tab = p->tab(B200P_SLOT_WEFF);
    a.old_mask = d_old_mask; a.new_mask = d_new_mask; a.st = p->d_state;
    a.mode = mode; a.force = force; a.forced_threshold = forced_threshold; a.outputs = outputs;
    bool vec = p->vec_ok[kslot];
    if (outputs & B200P_EMIT_MASKF) vec = vec && p->vec_ok[B200P_SLOT_MASKF];
    if (outputs & B200P_EMIT_WEFF) vec = vec && p->vec_ok[B200P_SLOT_WEFF] && p->vec_ok[B200P_SLOT_W];
    a.vec_ok = vec ? 1 : 0;
    if (patch && prov == d_new_mask) a.new_mask = d_new_mask;           // the sweep wrote straight into the destination
    else if (patch) a.new_mask = prov;                                     // full-pass fallback and patch both work on `prov`
    k_emit_masks<<<p->grid_for(chunk_end - chunk_begin, 4), kThreads, 0, st>>>(a, chunk_begin, chunk_end);
    B200P_LAUNCH_CHECK("k_emit_masks");
    if (patch && prov != d_new_mask)
        B200P_CUDA(cudaMemcpyAsync(d_new_mask + chunk_begin * kWordsPerChunk, prov + chunk_begin * kWordsPerChunk,
                                   (size_t)(chunk_end - chunk_begin) * kWordsPerChunk * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    return B200P_OK;
}

extern "C" int b200p_count_zeros(b200p_plan* p, const uint32_t* d_mask, uint64_t* d_out, int use_weights, void* stream) {
    B200P_REQUIRE(p != nullptr && d_out != nullptr, B200P_EINVAL, "count_zeros: null argument");
    B200P_REQUIRE(use_weights || d_mask, B200P_EINVAL, "count_zeros: need a mask or the weights");
    if (use_weights) B200P_REQUIRE(p->bound[B200P_SLOT_WEFF] || p->bound[B200P_SLOT_W], B200P_ESTATE, "count_zeros: W slot is not bound");
    B200P_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    B200P_CUDA(cudaMemsetAsync(d_out, 0, 2 * sizeof(uint64_t), st));
    k_count_zeros<<<p->grid_for(p->n_chunks, 4), kThreads, 0, st>>>(p->d_chunk_n, p->tab(B200P_SLOT_W), d_mask,
        (unsigned long long*)d_out, p->n_chunks, p->vec_ok[B200P_SLOT_W] ? 1 : 0, use_weights);
    B200P_LAUNCH_CHECK("k_count_zeros");
    return B200P_OK;
}

extern "C" int b200p_mask_pack_from_f32(b200p_plan* p, uint32_t* d_mask, void* stream) {
    B200P_REQUIRE(p != nullptr && d_mask != nullptr, B200P_EINVAL, "mask_pack: null argument");
    B200P_REQUIRE(p->bound[B200P_SLOT_MASKF], B200P_ESTATE, "mask_pack: MASKF slot is not bound");
    B200P_CUDA(cudaSetDevice(p->device));
    k_mask_convert<0><<<p->grid_for(p->n_chunks, 4), kThreads, 0, (cudaStream_t)stream>>>(p->d_chunk_n, p->tab(B200P_SLOT_MASKF),
        nullptr, nullptr, d_mask, p->n_chunks, p->vec_ok[B200P_SLOT_MASKF] ? 1 : 0, 0);
    B200P_LAUNCH_CHECK("k_mask_convert<0>");
    return B200P_OK;
}
extern "C" int b200p_mask_unpack_to_f32(b200p_plan* p, const uint32_t* d_mask, void* stream) {
    B200P_REQUIRE(p != nullptr && d_mask != nullptr, B200P_EINVAL, "mask_unpack: null argument");
    B200P_REQUIRE(p->bound[B200P_SLOT_MASKF], B200P_ESTATE, "mask_unpack: MASKF slot is not bound");
    B200P_CUDA(cudaSetDevice(p->device));
    k_mask_convert<1><<<p->grid_for(p->n_chunks, 4), kThreads, 0, (cudaStream_t)stream>>>(p->d_chunk_n, p->tab(B200P_SLOT_MASKF),
        nullptr, nullptr, const_cast<uint32_t*>(d_mask), p->n_chunks, p->vec_ok[B200P_SLOT_MASKF] ? 1 : 0, 0);
    B200P_LAUNCH_CHECK("k_mask_convert<1>");
    return B200P_OK;
}
extern "C" int b200p_apply_mask(b200p_plan* p, const uint32_t* d_mask, int outputs, void* stream) {
    B200P_REQUIRE(p != nullptr && d_mask != nullptr, B200P_EINVAL, "apply_mask: null argument");
    B200P_REQUIRE(p->bound[B200P_SLOT_W], B200P_ESTATE, "apply_mask: W slot is not bound");
    const bool want32 = outputs & B200P_EMIT_WEFF, want16 = outputs & B200P_SGD_EMIT_WEFF16;
    B200P_REQUIRE(want32 || want16, B200P_EINVAL, "apply_mask: no output requested");
    if (want32) B200P_REQUIRE(p->bound[B200P_SLOT_WEFF], B200P_ESTATE, "apply_mask: WEFF slot is not bound");
    if (want16) B200P_REQUIRE(p->bound[B200P_SLOT_WEFF16], B200P_ESTATE, "apply_mask: WEFF16 slot is not bound");
    B200P_CUDA(cudaSetDevice(p->device));
    bool vec = p->vec_ok[B200P_SLOT_W] && (!want32 || p->vec_ok[B200P_SLOT_WEFF]) && (!want16 || p->vec_ok[B200P_SLOT_WEFF16]);
    k_mask_convert<2><<<p->grid_for(p->n_chunks, 4), kThreads, 0, (cudaStream_t)stream>>>(p->d_chunk_n,
        want32 ? p->tab(B200P_SLOT_WEFF) : nullptr, p->tab(B200P_SLOT_W),
        want16 ? p->tab(B200P_SLOT_WEFF16) : nullptr, const_cast<uint32_t*>(d_mask), p->n_chunks,
        vec ? 1 : 0, want32 ? B200P_EMIT_WEFF : 0);
    B200P_LAUNCH_CHECK("k_mask_convert<2>");
    return B200P_OK;
}
extern "C" int b200p_mask_grads(b200p_plan* p, const uint32_t* d_mask, void* stream) {
    B200P_REQUIRE(p != nullptr && d_mask != nullptr, B200P_EINVAL, "mask_grads: null argument");
    B200P_REQUIRE(p->bound[B200P_SLOT_G], B200P_ESTATE, "mask_grads: G slot is not bound");
    B200P_CUDA(cudaSetDevice(p->device));
    k_mask_convert<3><<<p->grid_for(p->n_chunks, 4), kThreads, 0, (cudaStream_t)stream>>>(p->d_chunk_n, p->tab(B200P_SLOT_G),
        nullptr, nullptr, const_cast<uint32_t*>(d_mask), p->n_chunks, p->vec_ok[B200P_SLOT_G] ? 1 : 0, 0);
    B200P_LAUNCH_CHECK("k_mask_convert<3>");
    return B200P_OK;
}

// select + emit in one call: the sweep writes the provisional mask straight into d_new_mask, the emit patches it
extern "C" int b200p_mask_build(b200p_plan* p, int key_source, const uint32_t* d_old_mask, uint64_t k, int mode,
                                uint32_t* d_new_mask, void* stream) {
    B200P_REQUIRE(p != nullptr && d_new_mask != nullptr, B200P_EINVAL, "mask_build: null argument");
    B200P_REQUIRE(d_new_mask != d_old_mask, B200P_EINVAL, "mask_build: the new mask must not alias the old one");
    p->prov_target = d_new_mask;
    p->fuse_emit = true; p->emit_done = false;
    int rc = b200p_select_kth(p, key_source, d_old_mask, k, mode, stream);
    p->fuse_emit = false;
    if (rc) { p->prov_target = nullptr; return rc; }
    if (p->emit_done) {                       // the finish kernel patched the mask itself (one launch less, no barrier in between)
        p->emit_done = false; p->prov_armed = false; p->prov_target = nullptr;
        return B200P_OK;
    }
    return b200p_emit_masks(p, key_source, mode, 0, 0.f, d_old_mask, d_new_mask, 0, 0, -1, stream);
}
