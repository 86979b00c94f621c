// host.cu — host-buffer entry points: the complete mask build driven from HOST memory.
// This is the path bench.py reports as "e2e": host->device copies of the weights and of
// every mini-batch's gradients are inside the call, double-buffered against the score
// kernels on two streams, and the packed mask comes back to the host.
#include "common.cuh"
#include <math.h>

using namespace b200p;

static int ensure_arena(b200p_plan* p, bool need_grads) {
    B200P_CUDA(cudaSetDevice(p->device));
    const size_t nbytes = (size_t)p->total * sizeof(float);
    const size_t mbytes = (size_t)p->n_chunks * kWordsPerChunk * sizeof(uint32_t);
    if (!p->arena_w) B200P_CUDA(cudaMalloc(&p->arena_w, nbytes));
    if (!p->arena_mask) B200P_CUDA(cudaMalloc(&p->arena_mask, mbytes));
    if (!p->arena_old_mask) B200P_CUDA(cudaMalloc(&p->arena_old_mask, mbytes));
    if (need_grads) {
        for (int i = 0; i < 2; ++i) if (!p->arena_g[i]) B200P_CUDA(cudaMalloc(&p->arena_g[i], nbytes));
        if (!p->arena_score) B200P_CUDA(cudaMalloc(&p->arena_score, nbytes));
    }
    if (need_grads) {
        for (int i = 0; i < 2; ++i) if (!p->arena_gtab[i]) {
            std::vector<const void*> ptrs(p->n_seg);
            for (int t = 0; t < p->n_seg; ++t) ptrs[t] = p->arena_g[i] + p->seg_flat_start[t];
            int rc = b200p_ptrtable_create(p, B200P_SLOT_G, ptrs.data(), nullptr, &p->arena_gtab[i]);
            if (rc) return rc;
        }
        B200P_CUDA(cudaStreamSynchronize(nullptr));
    }
    for (int i = 0; i < 2; ++i) if (!p->arena_streams[i]) B200P_CUDA(cudaStreamCreateWithFlags(&p->arena_streams[i], cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) if (!p->arena_events[i]) B200P_CUDA(cudaEventCreateWithFlags(&p->arena_events[i], cudaEventDisableTiming));
    return B200P_OK;
}

// bind a flat device array (segment-concatenated, unpadded) to a slot
static int bind_flat(b200p_plan* p, int slot, float* base, cudaStream_t st) {
    std::vector<const void*> ptrs(p->n_seg);
    for (int t = 0; t < p->n_seg; ++t) ptrs[t] = base + p->seg_flat_start[t];
    return b200p_plan_bind(p, slot, ptrs.data(), st);
}

extern "C" int b200p_snip_mask_build_host(b200p_plan* p, const float* h_w, const float* const* h_g,
                                          int n_batches, uint64_t k, uint32_t* h_mask_out,
                                          b200p_select_result_t* h_result) {
    B200P_REQUIRE(p && h_w && h_g && h_mask_out, B200P_EINVAL, "snip_mask_build_host: null argument");
    B200P_REQUIRE(n_batches >= 1, B200P_EINVAL, "snip_mask_build_host: need at least one mini-batch");
    int rc = ensure_arena(p, true); if (rc) return rc;
    cudaStream_t copy = p->arena_streams[0], comp = p->arena_streams[1];
    cudaEvent_t* ev_copied = p->arena_events;        // [2] gradient buffer i is filled
    cudaEvent_t* ev_consumed = p->arena_events + 2;  // [2] gradient buffer i has been read
    const size_t nbytes = (size_t)p->total * sizeof(float);
    const size_t mbytes = (size_t)p->n_chunks * kWordsPerChunk * sizeof(uint32_t);

    rc = bind_flat(p, B200P_SLOT_W, p->arena_w, comp); if (rc) return rc;
    rc = bind_flat(p, B200P_SLOT_SCORE, p->arena_score, comp); if (rc) return rc;
    B200P_CUDA(cudaMemcpyAsync(p->arena_w, h_w, nbytes, cudaMemcpyHostToDevice, copy));
    for (int b = 0; b < n_batches; ++b) {
        const int i = b & 1;
        if (b >= 2) B200P_CUDA(cudaStreamWaitEvent(copy, ev_consumed[i], 0));
        B200P_CUDA(cudaMemcpyAsync(p->arena_g[i], h_g[b], nbytes, cudaMemcpyHostToDevice, copy));
        B200P_CUDA(cudaEventRecord(ev_copied[i], copy));
        B200P_CUDA(cudaStreamWaitEvent(comp, ev_copied[i], 0));
        rc = b200p_plan_bind_table(p, B200P_SLOT_G, p->arena_gtab[i]); if (rc) return rc;
        rc = b200p_score_accumulate(p, b > 0, 0, -1, comp); if (rc) return rc;
        B200P_CUDA(cudaEventRecord(ev_consumed[i], comp));
    }
    if (k >= (uint64_t)p->total) {          // train.py:300-301: threshold = +inf, prune everything
        rc = b200p_select_begin(p, 0, B200P_MODE_SNIP_STRICT, 0, comp); if (rc) return rc;
        rc = b200p_emit_masks(p, B200P_KEY_SCORE, B200P_MODE_SNIP_STRICT, 3, INFINITY, nullptr, p->arena_mask, 0, 0, -1, comp);
    } else if (k == 0) {                    // train.py:302-303: threshold = -1, keep everything
        rc = b200p_select_begin(p, 0, B200P_MODE_SNIP_STRICT, 0, comp); if (rc) return rc;
        rc = b200p_emit_masks(p, B200P_KEY_SCORE, B200P_MODE_SNIP_STRICT, 3, -1.0f, nullptr, p->arena_mask, 0, 0, -1, comp);
    } else {
        rc = b200p_mask_build(p, B200P_KEY_SCORE, nullptr, k, B200P_MODE_SNIP_STRICT, p->arena_mask, comp);      // sweep + patching emit
    }
    if (rc) return rc;
    B200P_CUDA(cudaMemcpyAsync(h_mask_out, p->arena_mask, mbytes, cudaMemcpyDeviceToHost, comp));
    if (h_result) { rc = b200p_select_result(p, h_result, comp); if (rc) return rc; }
    B200P_CUDA(cudaStreamSynchronize(comp));
    B200P_CUDA(cudaStreamSynchronize(copy));
    return B200P_OK;
}

extern "C" int b200p_magnitude_mask_build_host(b200p_plan* p, const float* h_w, const uint32_t* h_old_mask,
                                               uint64_t k, uint32_t* h_mask_out, b200p_select_result_t* h_result) {
    B200P_REQUIRE(p && h_w && h_mask_out, B200P_EINVAL, "magnitude_mask_build_host: null argument");
    int rc = ensure_arena(p, false); if (rc) return rc;
    cudaStream_t comp = p->arena_streams[1];
    const size_t nbytes = (size_t)p->total * sizeof(float);
    const size_t mbytes = (size_t)p->n_chunks * kWordsPerChunk * sizeof(uint32_t);
    rc = bind_flat(p, B200P_SLOT_W, p->arena_w, comp); if (rc) return rc;
    B200P_CUDA(cudaMemcpyAsync(p->arena_w, h_w, nbytes, cudaMemcpyHostToDevice, comp));
    const uint32_t* old_mask = nullptr;
    if (h_old_mask) {
        B200P_CUDA(cudaMemcpyAsync(p->arena_old_mask, h_old_mask, mbytes, cudaMemcpyHostToDevice, comp));
        old_mask = p->arena_old_mask;
    }
    if (k == 0) {                           // prune.py:533: nothing to prune, mask unchanged
        rc = b200p_select_begin(p, 0, B200P_MODE_EXACT_K, 0, comp); if (rc) return rc;
        rc = b200p_emit_masks(p, B200P_KEY_ABS_W, B200P_MODE_EXACT_K, 1, 0.f, old_mask, p->arena_mask, 0, 0, -1, comp);
    } else {
        rc = b200p_mask_build(p, B200P_KEY_ABS_W, old_mask, k, B200P_MODE_EXACT_K, p->arena_mask, comp);
    }
    if (rc) return rc;
    B200P_CUDA(cudaMemcpyAsync(h_mask_out, p->arena_mask, mbytes, cudaMemcpyDeviceToHost, comp));
    if (h_result) { rc = b200p_select_result(p, h_result, comp); if (rc) return rc; }
    B200P_CUDA(cudaStreamSynchronize(comp));
    return B200P_OK;
}
