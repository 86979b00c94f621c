// plan.cu — plan object: segment / chunk tables, pointer slots, workspace.
#include "common.cuh"
#include <string.h>
#include <new>

namespace b200p {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return B200P_ECUDA;
}
}  // namespace b200p

using namespace b200p;

extern "C" const char* b200p_last_error(void) { return g_last_error.c_str(); }
extern "C" int b200p_version(void) { return 100; }
extern "C" int b200p_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int b200p_plan_create(int device, int n_segments, const int64_t* h_numel,
                                 int64_t cand_capacity, b200p_plan** out) {
    B200P_REQUIRE(out != nullptr && h_numel != nullptr, B200P_EINVAL, "plan_create: null argument");
    B200P_REQUIRE(n_segments > 0, B200P_EINVAL, "plan_create: need at least one segment");
    *out = nullptr;
    B200P_CUDA(cudaSetDevice(device));
    b200p_plan* p = new (std::nothrow) b200p_plan();
    B200P_REQUIRE(p != nullptr, B200P_ENOMEM, "plan_create: out of host memory");
    p->device = device;
    p->n_seg = n_segments;
    p->numel.assign(h_numel, h_numel + n_segments);
    p->seg_chunk_start.resize(n_segments + 1);
    p->seg_flat_start.resize(n_segments + 1);
    int64_t chunks = 0, flat = 0;
    for (int t = 0; t < n_segments; ++t) {
        if (h_numel[t] <= 0) { delete p; set_error("plan_create: segment with numel <= 0"); return B200P_EINVAL; }
        p->seg_chunk_start[t] = chunks;
        p->seg_flat_start[t] = flat;
        chunks += (h_numel[t] + kChunk - 1) / kChunk;
        flat += h_numel[t];
    }
    p->seg_chunk_start[n_segments] = chunks;
    p->seg_flat_start[n_segments] = flat;
    p->n_chunks = chunks;
    p->total = flat;
    // candidate positions are 32-bit (chunk * 4096 + element)
    if (chunks * (int64_t)kChunk >= (int64_t)1 << 32) {
        delete p; set_error("plan_create: more than 2^32 padded elements are not supported"); return B200P_EINVAL;
    }
    if (cand_capacity <= 0) {
        cand_capacity = flat / 16;
        if (cand_capacity < (1 << 20)) cand_capacity = 1 << 20;
    }
    if (cand_capacity > flat) cand_capacity = flat;
    p->cand_capacity = cand_capacity;

    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaGetDeviceProperties"); }
    p->num_sms = prop.multiProcessorCount;

    std::vector<int32_t> chunk_seg(chunks);
    for (int t = 0; t < n_segments; ++t)
        for (int64_t c = p->seg_chunk_start[t]; c < p->seg_chunk_start[t + 1]; ++c) chunk_seg[c] = t;

#define TRY(call) do { e = (call); if (e != cudaSuccess) { b200p_plan_destroy(p); return cuda_fail(e, #call); } } while (0)
    TRY(cudaMalloc(&p->d_chunk_seg, chunks * sizeof(int32_t)));
    TRY(cudaMalloc(&p->d_seg_chunk_start, (n_segments + 1) * sizeof(int64_t)));
    TRY(cudaMalloc(&p->d_seg_numel, n_segments * sizeof(int64_t)));
    TRY(cudaMemcpy(p->d_chunk_seg, chunk_seg.data(), chunks * sizeof(int32_t), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(p->d_seg_chunk_start, p->seg_chunk_start.data(), (n_segments + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(p->d_seg_numel, p->numel.data(), n_segments * sizeof(int64_t), cudaMemcpyHostToDevice));
    for (int s = 0; s < B200P_NUM_SLOTS; ++s) TRY(cudaMalloc(&p->d_ptrs[s], n_segments * sizeof(void*)));
    TRY(cudaMalloc(&p->d_hist, kHistBins * sizeof(unsigned long long)));
    TRY(cudaMalloc(&p->d_state, sizeof(SelState)));
    TRY(cudaMalloc(&p->d_cand_key, cand_capacity * sizeof(uint32_t)));
    TRY(cudaMalloc(&p->d_cand_pos, cand_capacity * sizeof(uint32_t)));
    TRY(cudaMalloc(&p->d_chunk_ties, chunks * sizeof(uint32_t)));
    TRY(cudaMemset(p->d_hist, 0, kHistBins * sizeof(unsigned long long)));
    TRY(cudaMemset(p->d_state, 0, sizeof(SelState)));
    TRY(cudaMemset(p->d_chunk_ties, 0, chunks * sizeof(uint32_t)));
    TRY(cudaDeviceSynchronize());
#undef TRY
    *out = p;
    return B200P_OK;
}

extern "C" int b200p_plan_destroy(b200p_plan* p) {
    if (!p) return B200P_OK;
    cudaSetDevice(p->device);
    cudaFree(p->d_chunk_seg); cudaFree(p->d_seg_chunk_start); cudaFree(p->d_seg_numel);
    for (int s = 0; s < B200P_NUM_SLOTS; ++s) cudaFree(p->d_ptrs[s]);
    cudaFree(p->d_hist); cudaFree(p->d_state); cudaFree(p->d_cand_key); cudaFree(p->d_cand_pos);
    cudaFree(p->d_chunk_ties);
    cudaFree(p->arena_w); cudaFree(p->arena_g[0]); cudaFree(p->arena_g[1]); cudaFree(p->arena_score);
    cudaFree(p->arena_mask); cudaFree(p->arena_old_mask);
    for (int i = 0; i < 2; ++i) if (p->arena_streams[i]) cudaStreamDestroy(p->arena_streams[i]);
    for (int i = 0; i < 4; ++i) if (p->arena_events[i]) cudaEventDestroy(p->arena_events[i]);
    cudaGetLastError();
    delete p;
    return B200P_OK;
}

extern "C" int64_t b200p_plan_total(const b200p_plan* p) { return p ? p->total : -1; }
extern "C" int64_t b200p_plan_num_chunks(const b200p_plan* p) { return p ? p->n_chunks : -1; }
extern "C" int64_t b200p_plan_mask_words(const b200p_plan* p) { return p ? p->n_chunks * kWordsPerChunk : -1; }
extern "C" int64_t b200p_plan_seg_chunk_start(const b200p_plan* p, int seg) {
    if (!p || seg < 0 || seg > p->n_seg) return -1;
    return p->seg_chunk_start[seg];
}
extern "C" int64_t b200p_plan_chunk_flat_start(const b200p_plan* p, int64_t chunk) {
    if (!p || chunk < 0 || chunk > p->n_chunks) return -1;
    if (chunk == p->n_chunks) return p->total;
    // segment that owns the chunk: last t with seg_chunk_start[t] <= chunk
    int lo = 0, hi = p->n_seg - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) / 2; if (p->seg_chunk_start[mid] <= chunk) lo = mid; else hi = mid - 1; }
    return p->seg_flat_start[lo] + (chunk - p->seg_chunk_start[lo]) * (int64_t)kChunk;
}
extern "C" void* b200p_plan_hist_ptr(b200p_plan* p) { return p ? (void*)p->d_hist : nullptr; }
extern "C" void* b200p_plan_state_ptr(b200p_plan* p) { return p ? (void*)p->d_state : nullptr; }

namespace b200p {
// pointer tables travel as kernel arguments (no staging buffer, no host sync): 256 per launch
struct PtrPack { void* p[256]; };
__global__ void k_set_ptrs(void** dst, PtrPack pack, int n) {
    if ((int)threadIdx.x < n) dst[threadIdx.x] = pack.p[threadIdx.x];
}
}  // namespace b200p

extern "C" int b200p_plan_bind(b200p_plan* p, int slot, const void* const* h_ptrs, void* stream) {
    B200P_REQUIRE(p != nullptr && h_ptrs != nullptr, B200P_EINVAL, "plan_bind: null argument");
    B200P_REQUIRE(slot >= 0 && slot < B200P_NUM_SLOTS, B200P_EINVAL, "plan_bind: bad slot");
    B200P_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    bool vec = true;
    for (int t0 = 0; t0 < p->n_seg; t0 += 256) {
        PtrPack pack;
        const int n = p->n_seg - t0 < 256 ? p->n_seg - t0 : 256;
        for (int i = 0; i < n; ++i) {
            B200P_REQUIRE(h_ptrs[t0 + i] != nullptr, B200P_EINVAL, "plan_bind: null segment pointer");
            pack.p[i] = const_cast<void*>(h_ptrs[t0 + i]);
            if ((uintptr_t)h_ptrs[t0 + i] & 15u) vec = false;
        }
        k_set_ptrs<<<1, 256, 0, st>>>(p->d_ptrs[slot] + t0, pack, n);
        B200P_LAUNCH_CHECK("k_set_ptrs");
    }
    p->bound[slot] = true;
    p->vec_ok[slot] = vec;
    return B200P_OK;
}
