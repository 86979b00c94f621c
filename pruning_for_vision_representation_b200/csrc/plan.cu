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
extern "C" int b200p_version(void) { return 200; }
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

    std::vector<int32_t> chunk_seg(chunks), chunk_n(chunks);
    std::vector<int64_t> chunk_elem0(chunks);
    for (int t = 0; t < n_segments; ++t)
        for (int64_t c = p->seg_chunk_start[t]; c < p->seg_chunk_start[t + 1]; ++c) {
            const int64_t e0 = (c - p->seg_chunk_start[t]) * (int64_t)kChunk;
            const int64_t rem = h_numel[t] - e0;
            chunk_seg[c] = t; chunk_elem0[c] = e0; chunk_n[c] = rem < kChunk ? (int32_t)rem : kChunk;
        }

#define TRY(call) do { e = (call); if (e != cudaSuccess) { b200p_plan_destroy(p); return cuda_fail(e, #call); } } while (0)
    TRY(cudaMalloc(&p->d_chunk_seg, chunks * sizeof(int32_t)));
    TRY(cudaMalloc(&p->d_chunk_n, chunks * sizeof(int32_t)));
    TRY(cudaMalloc(&p->d_chunk_elem0, chunks * sizeof(int64_t)));
    TRY(cudaMemcpy(p->d_chunk_seg, chunk_seg.data(), chunks * sizeof(int32_t), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(p->d_chunk_n, chunk_n.data(), chunks * sizeof(int32_t), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(p->d_chunk_elem0, chunk_elem0.data(), chunks * sizeof(int64_t), cudaMemcpyHostToDevice));
    for (int s = 0; s < B200P_NUM_SLOTS; ++s) TRY(cudaMalloc(&p->d_tab_own[s], chunks * sizeof(void*)));
    TRY(cudaMalloc(&p->d_hist, (size_t)kHistReplicas * kHistStride * sizeof(unsigned long long)));
    TRY(cudaMalloc(&p->d_state, sizeof(SelState)));
    TRY(cudaMalloc(&p->d_cand_key, cand_capacity * sizeof(uint32_t)));
    TRY(cudaMalloc(&p->d_cand_pos, cand_capacity * sizeof(uint32_t)));
    TRY(cudaMalloc(&p->d_chunk_ties, (chunks + kTieListCap + 8) * sizeof(uint32_t)));      // table + [count, positions...] of the tie list
    TRY(cudaMalloc(&p->d_prov, chunks * kWordsPerChunk * sizeof(uint32_t)));
    TRY(cudaMalloc(&p->d_rank_ties, 8 * sizeof(unsigned long long)));
    TRY(cudaMalloc(&p->d_sample_cache, (kHistBins + kHistExtra) * sizeof(unsigned long long)));
    TRY(cudaMemset(p->d_rank_ties, 0, 8 * sizeof(unsigned long long)));
    TRY(cudaMemset(p->d_hist, 0, (size_t)kHistReplicas * kHistStride * sizeof(unsigned long long)));
    TRY(cudaMemset(p->d_state, 0, sizeof(SelState)));
    TRY(cudaMemset(p->d_chunk_ties, 0, (chunks + kTieListCap + 8) * sizeof(uint32_t)));
    TRY(cudaDeviceSynchronize());
#undef TRY
    *out = p;
    return B200P_OK;
}

extern "C" int b200p_plan_destroy(b200p_plan* p) {
    if (!p) return B200P_OK;
    cudaSetDevice(p->device);
    cudaFree(p->d_chunk_seg); cudaFree(p->d_chunk_n); cudaFree(p->d_chunk_elem0);
    for (int s = 0; s < B200P_NUM_SLOTS; ++s) cudaFree(p->d_tab_own[s]);
    cudaFree(p->d_hist); cudaFree(p->d_state); cudaFree(p->d_cand_key); cudaFree(p->d_cand_pos);
    cudaFree(p->d_chunk_ties); cudaFree(p->d_prov); cudaFree(p->d_rank_ties); cudaFree(p->d_sample_cache);
    for (int i = 0; i < 2; ++i) if (p->arena_gtab[i]) { b200p_ptrtable_destroy(p->arena_gtab[i]); p->arena_gtab[i] = nullptr; }
    cudaFree(p->arena_w); cudaFree(p->arena_g[0]); cudaFree(p->arena_g[1]); cudaFree(p->arena_score);
    cudaFree(p->arena_mask); cudaFree(p->arena_old_mask);
    for (int i = 0; i < 2; ++i) if (p->arena_streams[i]) cudaStreamDestroy(p->arena_streams[i]);
    for (int i = 0; i < 4; ++i) if (p->arena_events[i]) cudaEventDestroy(p->arena_events[i]);
    for (auto& e : p->tev) cudaEventDestroy(e);
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
extern "C" int b200p_plan_set_option(b200p_plan* p, int option, int64_t value) {
    B200P_REQUIRE(p != nullptr, B200P_EINVAL, "plan_set_option: null plan");
    if (option == B200P_OPT_SELECT_IMPL) {
        B200P_REQUIRE(value == B200P_SELECT_SAMPLED || value == B200P_SELECT_EXACT, B200P_EINVAL, "plan_set_option: bad select implementation");
        p->select_impl = (int)value;
        return B200P_OK;
    }
    if (option == B200P_OPT_REUSE_SAMPLE) {
        p->reuse_sample = value != 0;
        if (!p->reuse_sample) p->sample_cache_valid = false;
        return B200P_OK;
    }
    if (option == B200P_OPT_COOP_GRID) {
        B200P_REQUIRE(value >= 0 && value <= 4096, B200P_EINVAL, "plan_set_option: bad cooperative grid limit");
        p->coop_grid_limit = (int)value;
        return B200P_OK;
    }
    if (option == B200P_OPT_TIME_SWEEP) {
        p->time_sweep = value != 0;
        return B200P_OK;
    }
    set_error("plan_set_option: unknown option");
    return B200P_EINVAL;
}

namespace b200p {
constexpr int kTimedPairs = 64;
static int fold_oldest_pair(b200p_plan* p) {
    const int idx = ((p->tev_head - p->tev_pending) % kTimedPairs + kTimedPairs) % kTimedPairs;
    B200P_CUDA(cudaEventSynchronize(p->tev[2 * idx + 1]));
    float ms = 0.f;
    B200P_CUDA(cudaEventElapsedTime(&ms, p->tev[2 * idx], p->tev[2 * idx + 1]));
    p->tev_sum_ms += ms; p->tev_n += 1; p->tev_pending -= 1;
    return B200P_OK;
}
// called by the fused SNIP sequence right before (which = 0) and after (which = 1) the sweep launch
int plan_time_mark(b200p_plan* p, int which, cudaStream_t st) {
    if (!p->time_sweep) return B200P_OK;
    if (p->tev.empty()) {
        p->tev.resize(2 * kTimedPairs);
        for (auto& e : p->tev) B200P_CUDA(cudaEventCreate(&e));
    }
    if (which == 0 && p->tev_pending == kTimedPairs) { int rc = fold_oldest_pair(p); if (rc) return rc; }
    B200P_CUDA(cudaEventRecord(p->tev[2 * p->tev_head + which], st));
    if (which == 1) { p->tev_head = (p->tev_head + 1) % kTimedPairs; p->tev_pending += 1; }
    return B200P_OK;
}
}  // namespace b200p

extern "C" int b200p_plan_kernel_time_ms(b200p_plan* p, double* out_ms, int64_t* out_launches) {
    B200P_REQUIRE(p != nullptr && out_ms != nullptr && out_launches != nullptr, B200P_EINVAL, "plan_kernel_time_ms: null argument");
    B200P_CUDA(cudaSetDevice(p->device));
    while (p->tev_pending > 0) { int rc = fold_oldest_pair(p); if (rc) return rc; }
    *out_launches = p->tev_n;
    *out_ms = p->tev_n ? p->tev_sum_ms / (double)p->tev_n : 0.0;
    p->tev_sum_ms = 0.0; p->tev_n = 0;
    return B200P_OK;
}
extern "C" void* b200p_plan_hist_ptr(b200p_plan* p) { return p ? (void*)p->d_hist : nullptr; }
extern "C" void* b200p_plan_state_ptr(b200p_plan* p) { return p ? (void*)p->d_state : nullptr; }

namespace b200p {
// Segment pointers travel as kernel arguments (no staging buffer, no host sync), 256 per launch;
// the kernel expands them into the per-chunk table: tab[c] = seg_ptr[chunk_seg[c]] + elem0[c]*size.
struct PtrPack { void* p[256]; };
__global__ void k_fill_chunk_ptrs(void** __restrict__ tab, const int32_t* __restrict__ chunk_seg,
                                  const int64_t* __restrict__ chunk_elem0, PtrPack pack, int seg0, int nseg,
                                  int elem_size, int64_t n_chunks) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const int t = chunk_seg[c] - seg0;
    if (t < 0 || t >= nseg) return;
    tab[c] = (char*)pack.p[t] + chunk_elem0[c] * elem_size;
}

static int slot_elem_size(int slot) { return slot == B200P_SLOT_WEFF16 ? 2 : 4; }

static int fill_table(b200p_plan* p, void** d_tab, const void* const* h_ptrs, int elem_size, cudaStream_t st, bool* vec_out) {
    bool vec = true;
    for (int t0 = 0; t0 < p->n_seg; t0 += 256) {
        PtrPack pack;
        const int n = p->n_seg - t0 < 256 ? p->n_seg - t0 : 256;
        for (int i = 0; i < n; ++i) {
            if (h_ptrs[t0 + i] == nullptr) { set_error("bind: null segment pointer"); return B200P_EINVAL; }
            pack.p[i] = const_cast<void*>(h_ptrs[t0 + i]);
            if ((uintptr_t)h_ptrs[t0 + i] & 15u) vec = false;
        }
        // only the chunks of segments [t0, t0+n) are touched by this launch
        const int64_t c0 = p->seg_chunk_start[t0], c1 = p->seg_chunk_start[t0 + n];
        const int threads = 256;
        const int blocks = (int)((c1 - c0 + threads - 1) / threads);
        k_fill_chunk_ptrs<<<blocks, threads, 0, st>>>(d_tab + c0, p->d_chunk_seg + c0, p->d_chunk_elem0 + c0, pack, t0, n,
                                                     elem_size, c1 - c0);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "k_fill_chunk_ptrs");
    }
    *vec_out = vec;
    return B200P_OK;
}
}  // namespace b200p

extern "C" int b200p_plan_bind(b200p_plan* p, int slot, const void* const* h_ptrs, void* stream) {
    B200P_REQUIRE(p != nullptr && h_ptrs != nullptr, B200P_EINVAL, "plan_bind: null argument");
    B200P_REQUIRE(slot >= 0 && slot < B200P_NUM_SLOTS, B200P_EINVAL, "plan_bind: bad slot");
    B200P_CUDA(cudaSetDevice(p->device));
    bool vec = false;
    int rc = fill_table(p, p->d_tab_own[slot], h_ptrs, slot_elem_size(slot), (cudaStream_t)stream, &vec);
    if (rc) return rc;
    p->d_tab[slot] = p->d_tab_own[slot];
    p->bound[slot] = true;
    p->vec_ok[slot] = vec;
    if (slot == B200P_SLOT_W || slot == B200P_SLOT_SCORE) p->sample_cache_valid = false;      // other data behind the keys
    return B200P_OK;
}

extern "C" int b200p_ptrtable_create(b200p_plan* p, int slot, const void* const* h_ptrs, void* stream,
                                     b200p_ptrtable** out) {
    B200P_REQUIRE(p != nullptr && h_ptrs != nullptr && out != nullptr, B200P_EINVAL, "ptrtable_create: null argument");
    B200P_REQUIRE(slot >= 0 && slot < B200P_NUM_SLOTS, B200P_EINVAL, "ptrtable_create: bad slot");
    *out = nullptr;
    B200P_CUDA(cudaSetDevice(p->device));
    b200p_ptrtable* t = new (std::nothrow) b200p_ptrtable();
    B200P_REQUIRE(t != nullptr, B200P_ENOMEM, "ptrtable_create: out of host memory");
    t->plan = p;
    cudaError_t e = cudaMalloc(&t->d_tab, p->n_chunks * sizeof(void*));
    if (e != cudaSuccess) { delete t; return cuda_fail(e, "cudaMalloc(ptrtable)"); }
    int rc = fill_table(p, t->d_tab, h_ptrs, slot_elem_size(slot), (cudaStream_t)stream, &t->vec_ok);
    if (rc) { cudaFree(t->d_tab); delete t; return rc; }
    *out = t;
    return B200P_OK;
}

namespace b200p {
// several tables in one launch: up to 1024 segment pointers travel as kernel arguments
struct TabPack { void** tab[kMultiTabs]; };
__global__ void k_fill_chunk_ptrs_multi(TabPack tabs, const int32_t* __restrict__ chunk_seg, const int64_t* __restrict__ chunk_elem0,
                                        PtrPackBig pack, int n_seg, int n_tabs, int elem_size, int64_t n_chunks) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chunks * n_tabs) return;
    const int t = (int)(i / n_chunks);
    const int64_t c = i - (int64_t)t * n_chunks;
    tabs.tab[t][c] = (char*)pack.p[t * n_seg + chunk_seg[c]] + chunk_elem0[c] * elem_size;
}
}  // namespace b200p

extern "C" int b200p_ptrtables_update(b200p_ptrtable* const* tables, const void* const* const* h_ptrs, int n_tables, int slot, void* stream) {
    B200P_REQUIRE(tables != nullptr && h_ptrs != nullptr && n_tables >= 1, B200P_EINVAL, "ptrtables_update: bad argument");
    B200P_REQUIRE(slot >= 0 && slot < B200P_NUM_SLOTS, B200P_EINVAL, "ptrtables_update: bad slot");
    b200p_plan* p = tables[0] ? tables[0]->plan : nullptr;
    B200P_REQUIRE(p != nullptr, B200P_EINVAL, "ptrtables_update: null table");
    for (int i = 0; i < n_tables; ++i)
        B200P_REQUIRE(tables[i] != nullptr && tables[i]->plan == p && h_ptrs[i] != nullptr, B200P_EINVAL, "ptrtables_update: tables must belong to one plan");
    B200P_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int elem = slot_elem_size(slot);
    if (p->n_seg > kMultiPtrs) {                       // more segments than one launch carries: table by table
        for (int i = 0; i < n_tables; ++i) { int rc = fill_table(p, tables[i]->d_tab, h_ptrs[i], elem, st, &tables[i]->vec_ok); if (rc) return rc; }
        return B200P_OK;
    }
    const int per = kMultiPtrs / p->n_seg < kMultiTabs ? kMultiPtrs / p->n_seg : kMultiTabs;
    for (int i0 = 0; i0 < n_tables; i0 += per) {
        const int nt = n_tables - i0 < per ? n_tables - i0 : per;
        PtrPackBig pack; TabPack tabs;
        for (int i = 0; i < nt; ++i) {
            bool vec = true;
            tabs.tab[i] = tables[i0 + i]->d_tab;
            for (int t = 0; t < p->n_seg; ++t) {
                const void* q = h_ptrs[i0 + i][t];
                B200P_REQUIRE(q != nullptr, B200P_EINVAL, "ptrtables_update: null segment pointer");
                pack.p[i * p->n_seg + t] = const_cast<void*>(q);
                if ((uintptr_t)q & 15u) vec = false;
            }
            tables[i0 + i]->vec_ok = vec;
        }
        const int64_t work = p->n_chunks * nt;
        k_fill_chunk_ptrs_multi<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(tabs, p->d_chunk_seg, p->d_chunk_elem0, pack, p->n_seg, nt, elem, p->n_chunks);
        B200P_LAUNCH_CHECK("k_fill_chunk_ptrs_multi");
    }
    // a table currently bound to a slot keeps its binding; refresh the slot's vector flag
    for (int i = 0; i < n_tables; ++i)
        for (int sl = 0; sl < B200P_NUM_SLOTS; ++sl)
            if (p->d_tab[sl] == tables[i]->d_tab) p->vec_ok[sl] = tables[i]->vec_ok;
    return B200P_OK;
}

extern "C" int b200p_ptrtable_destroy(b200p_ptrtable* t) {
    if (!t) return B200P_OK;
    if (t->plan) {
        cudaSetDevice(t->plan->device);
        for (int s = 0; s < B200P_NUM_SLOTS; ++s)
            if (t->plan->d_tab[s] == t->d_tab) { t->plan->d_tab[s] = nullptr; t->plan->bound[s] = false; }
    }
    cudaFree(t->d_tab);
    cudaGetLastError();
    delete t;
    return B200P_OK;
}

extern "C" int b200p_plan_bind_table(b200p_plan* p, int slot, const b200p_ptrtable* t) {
    B200P_REQUIRE(p != nullptr && t != nullptr, B200P_EINVAL, "plan_bind_table: null argument");
    B200P_REQUIRE(slot >= 0 && slot < B200P_NUM_SLOTS, B200P_EINVAL, "plan_bind_table: bad slot");
    B200P_REQUIRE(t->plan == p, B200P_EINVAL, "plan_bind_table: table belongs to another plan");
    p->d_tab[slot] = t->d_tab;          // host-side swap: no launch, takes effect for the next launch
    p->bound[slot] = true;
    p->vec_ok[slot] = t->vec_ok;
    return B200P_OK;
}
