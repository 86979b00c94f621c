// comm.cu — host side of the peer-memory window (comm.cuh) and the bulk exchange kernels of the sharded mask build:
// mask-word push (all-gather of the packed mask), score-table construction for the fused score + scatter, barrier.
#include "common.cuh"
#include "comm.cuh"
#include <new>
#include <string.h>

namespace b200p {

static long long align256(long long v) { return (v + 255) / 256 * 256; }

static CommLayout make_layout(int world, long long mask_words, long long score_cap) {
    CommLayout l;
    long long o = 0;
    l.flags = o;      o = align256(o + (long long)CH_COUNT * kCommMaxWorld * 4);
    l.err = o;        o = align256(o + 4);
    l.trace = o;      o = align256(o + (long long)CH_COUNT * 2 * 8 * 8);
    l.hist_bins = o;  o = align256(o + 2ll * world * kCommHistBins * 4);
    l.hist_extra = o; o = align256(o + 2ll * world * kCommHistExtra * 8);
    l.gather = o;     o = align256(o + 2ll * world * kCommGatherWords * 4);
    l.mask = o;       o = align256(o + mask_words * 4);
    l.score = o;      o = align256(o + (long long)world * score_cap * 4);
    l.total = o;
    return l;
}

// ---- all-gather of the packed mask: every rank pushes the words of its chunk range into every window ----------------
__global__ void __launch_bounds__(256)
k_mask_push(CommDev c, long long w_begin, long long w_end, uint32_t seq, unsigned int* ticket) {
    comm_push_mask_words(c, w_begin, w_end);
    // last CTA: everybody's stores are out (fence + ticket), tell the peers and wait for theirs
    __shared__ unsigned int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    if (threadIdx.x == 0) *ticket = 0u;
    comm_signal_and_wait(c, CH_MASK, seq);
}

__global__ void k_comm_barrier(CommDev c, uint32_t seq) { comm_signal_and_wait(c, CH_BARRIER, seq); }

// per-chunk destination of the fused score + scatter: chunk ch is owned by rank o (chunk ranges in `bounds`), its partial
// scores go to part `rank` of o's score area at the chunk's offset inside o's slice
__global__ void k_fill_score_push_table(void** __restrict__ tab, CommDev c, const long long* __restrict__ chunk_flat, const long long* __restrict__ bounds,
                                        long long n_chunks) {
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n_chunks) return;
    int o = 0;
    while (o + 1 < c.world && ch >= bounds[o + 1]) ++o;
    const long long off = chunk_flat[ch] - chunk_flat[bounds[o]];
    tab[ch] = reinterpret_cast<float*>(c.win[o] + c.lay.score) + (long long)c.rank * c.score_cap + off;
}

}  // namespace b200p

using namespace b200p;

extern "C" int b200p_comm_create(int device, int rank, int world, int64_t mask_words, int64_t score_cap, b200p_comm** out) {
    B200P_REQUIRE(out != nullptr, B200P_EINVAL, "comm_create: null argument");
    B200P_REQUIRE(world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world, B200P_EINVAL, "comm_create: need 1 <= world <= 8, 0 <= rank < world");
    B200P_REQUIRE(mask_words >= 0 && score_cap >= 0, B200P_EINVAL, "comm_create: negative size");
    *out = nullptr;
    B200P_CUDA(cudaSetDevice(device));
    b200p_comm* c = new (std::nothrow) b200p_comm();
    B200P_REQUIRE(c != nullptr, B200P_ENOMEM, "comm_create: out of host memory");
    c->device = device; c->rank = rank; c->world = world; c->mask_words = mask_words; c->score_cap = (score_cap + 3) / 4 * 4;
    c->lay = make_layout(world, mask_words, c->score_cap);
    cudaError_t e = cudaMalloc(&c->window, (size_t)c->lay.total);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaMalloc(comm window)"); }
    e = cudaMemset(c->window, 0, (size_t)c->lay.total);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(c->window); delete c; return cuda_fail(e, "cudaMemset(comm window)"); }
    c->peers[rank] = c->window;
    if (world == 1) c->connected = true;
    *out = c;
    return B200P_OK;
}

extern "C" int b200p_comm_destroy(b200p_comm* c) {
    if (!c) return B200P_OK;
    cudaSetDevice(c->device);
    for (int p = 0; p < c->world; ++p)
        if (c->opened_ipc[p] && c->peers[p]) cudaIpcCloseMemHandle(c->peers[p]);
    cudaFree(c->window);
    cudaGetLastError();
    delete c;
    return B200P_OK;
}

extern "C" int b200p_comm_trace(b200p_comm* c, uint64_t* h_out64) {
    B200P_REQUIRE(c != nullptr && h_out64 != nullptr, B200P_EINVAL, "comm_trace: null argument");
    B200P_CUDA(cudaSetDevice(c->device));
    B200P_CUDA(cudaDeviceSynchronize());
    B200P_CUDA(cudaMemcpy(h_out64, c->window + c->lay.trace, (size_t)CH_COUNT * 2 * 8 * 8, cudaMemcpyDeviceToHost));
    return B200P_OK;
}

extern "C" int64_t b200p_comm_window_bytes(const b200p_comm* c) { return c ? c->lay.total : -1; }
extern "C" void* b200p_comm_window(const b200p_comm* c) { return c ? (void*)c->window : nullptr; }
extern "C" void* b200p_comm_mask_ptr(const b200p_comm* c) { return c ? (void*)(c->window + c->lay.mask) : nullptr; }

extern "C" int b200p_comm_ipc_handle(b200p_comm* c, void* h_out64) {
    B200P_REQUIRE(c != nullptr && h_out64 != nullptr, B200P_EINVAL, "comm_ipc_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    B200P_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    B200P_CUDA(cudaIpcGetMemHandle(&h, c->window));
    memcpy(h_out64, &h, 64);
    return B200P_OK;
}

extern "C" int b200p_comm_connect_ipc(b200p_comm* c, const void* h_handles) {
    B200P_REQUIRE(c != nullptr && h_handles != nullptr, B200P_EINVAL, "comm_connect_ipc: null argument");
    B200P_CUDA(cudaSetDevice(c->device));
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)h_handles + 64 * p, 64);
        void* ptr = nullptr;
        B200P_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peers[p] = (char*)ptr; c->opened_ipc[p] = true;
    }
    c->connected = true;
    return B200P_OK;
}

extern "C" int b200p_comm_connect_local(b200p_comm* c, void* const* h_peer_windows) {
    B200P_REQUIRE(c != nullptr && h_peer_windows != nullptr, B200P_EINVAL, "comm_connect_local: null argument");
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) continue;
        B200P_REQUIRE(h_peer_windows[p] != nullptr, B200P_EINVAL, "comm_connect_local: null peer window");
        c->peers[p] = (char*)h_peer_windows[p];
    }
    c->connected = true;
    return B200P_OK;
}

// 0 = fine; ch + 1 = a wait on channel ch timed out (a peer never arrived).  Synchronises the stream.
extern "C" int b200p_comm_error(b200p_comm* c, void* stream) {
    B200P_REQUIRE(c != nullptr, B200P_EINVAL, "comm_error: null argument");
    B200P_CUDA(cudaSetDevice(c->device));
    uint32_t v = 0;
    B200P_CUDA(cudaMemcpyAsync(&v, c->window + c->lay.err, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    B200P_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return (int)v;
}

extern "C" int b200p_comm_barrier(b200p_comm* c, void* stream) {
    B200P_REQUIRE(c != nullptr && c->connected, B200P_ESTATE, "comm_barrier: not connected");
    B200P_CUDA(cudaSetDevice(c->device));
    k_comm_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(c->dev(), ++c->seq[CH_BARRIER]);
    B200P_LAUNCH_CHECK("k_comm_barrier");
    return B200P_OK;
}

extern "C" int b200p_comm_mask_allgather(b200p_comm* c, b200p_plan* p, int64_t chunk_begin, int64_t chunk_end, void* stream) {
    B200P_REQUIRE(c != nullptr && p != nullptr && c->connected, B200P_ESTATE, "comm_mask_allgather: not connected");
    B200P_REQUIRE(chunk_begin >= 0 && chunk_begin <= chunk_end && chunk_end <= p->n_chunks && p->n_chunks * kWordsPerChunk == c->mask_words,
                  B200P_EINVAL, "comm_mask_allgather: bad chunk range / mask size");
    B200P_CUDA(cudaSetDevice(c->device));
    const long long w0 = chunk_begin * kWordsPerChunk, w1 = chunk_end * kWordsPerChunk;
    long long blocks = ((w1 - w0) / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 4ll * p->num_sms) blocks = 4ll * p->num_sms;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(p->d_hist + kHistBins + kHistExtra - 1);     // last extra slot: unused by the select
    k_mask_push<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(c->dev(), w0, w1, ++c->seq[CH_MASK], ticket);
    B200P_LAUNCH_CHECK("k_mask_push");
    return B200P_OK;
}

// Pointer table for the SCORE slot whose entries point into the OWNERS' score areas: binding it makes the score kernels
// (b200p_score_accumulate / _multi) write each chunk's partial scores straight into the window of the rank that owns the
// chunk — the local score pass and the all-to-all of SURVEY §8(e) are one kernel.  h_bounds: world + 1 chunk bounds.
extern "C" int b200p_comm_score_push_table(b200p_comm* c, b200p_plan* p, const int64_t* h_bounds, void* stream, b200p_ptrtable** out) {
    B200P_REQUIRE(c != nullptr && p != nullptr && h_bounds != nullptr && out != nullptr && c->connected, B200P_ESTATE, "comm_score_push_table: bad argument / not connected");
    *out = nullptr;
    B200P_CUDA(cudaSetDevice(c->device));
    std::vector<long long> flat(p->n_chunks + 1), bounds(c->world + 1);
    for (int64_t ch = 0; ch <= p->n_chunks; ++ch) flat[ch] = b200p_plan_chunk_flat_start(p, ch);
    bool vec = true;
    for (int r = 0; r <= c->world; ++r) {
        bounds[r] = h_bounds[r];
        B200P_REQUIRE(bounds[r] >= 0 && bounds[r] <= p->n_chunks && (r == 0 || bounds[r] >= bounds[r - 1]), B200P_EINVAL, "comm_score_push_table: bad bounds");
    }
    for (int r = 0; r < c->world; ++r)
        B200P_REQUIRE(flat[bounds[r + 1]] - flat[bounds[r]] <= c->score_cap, B200P_EINVAL, "comm_score_push_table: a slice exceeds the score area");
    for (int64_t ch = 0; ch < p->n_chunks; ++ch) if (flat[ch] & 3) vec = false;          // 16-byte aligned destinations
    long long *d_flat = nullptr, *d_bounds = nullptr;
    B200P_CUDA(cudaMalloc(&d_flat, (p->n_chunks + 1) * sizeof(long long)));
    B200P_CUDA(cudaMalloc(&d_bounds, (c->world + 1) * sizeof(long long)));
    B200P_CUDA(cudaMemcpy(d_flat, flat.data(), (p->n_chunks + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    B200P_CUDA(cudaMemcpy(d_bounds, bounds.data(), (c->world + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    b200p_ptrtable* t = new (std::nothrow) b200p_ptrtable();
    B200P_REQUIRE(t != nullptr, B200P_ENOMEM, "comm_score_push_table: out of host memory");
    t->plan = p; t->vec_ok = vec;
    cudaError_t e = cudaMalloc(&t->d_tab, p->n_chunks * sizeof(void*));
    if (e != cudaSuccess) { delete t; cudaFree(d_flat); cudaFree(d_bounds); return cuda_fail(e, "cudaMalloc(score push table)"); }
    k_fill_score_push_table<<<(unsigned)((p->n_chunks + 255) / 256), 256, 0, (cudaStream_t)stream>>>(t->d_tab, c->dev(), d_flat, d_bounds, p->n_chunks);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(d_flat); cudaFree(d_bounds);
    if (e != cudaSuccess) { cudaFree(t->d_tab); delete t; return cuda_fail(e, "k_fill_score_push_table"); }
    *out = t;
    return B200P_OK;
}

// device address of this rank's score area (world parts of score_cap floats): the input of b200p_sum_parts
extern "C" void* b200p_comm_score_ptr(const b200p_comm* c) { return c ? (void*)(c->window + c->lay.score) : nullptr; }
extern "C" int64_t b200p_comm_score_cap(const b200p_comm* c) { return c ? c->score_cap : -1; }
