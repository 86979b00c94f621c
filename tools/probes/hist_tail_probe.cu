// Where do the ~10 us "last CTA" tails of the sample / sweep kernels go?  (ncu: sm__cycles_active.max - .min = 25 k cycles.)
// Mimics the end of k_select_bracket: every CTA flushes a 4096-bin shared histogram into one global histogram with
// atomics, takes a ticket, the last CTA reads the bins back, scans, clears.  globaltimer stamps per phase.
//   variant 0: u64 atomics, one global histogram (what select.cu does)
//   variant 1: u32 atomics
//   variant 2: cluster of 8, distributed DSMEM reduce (each CTA sums 1/8 of the bins over the 8 CTAs), then u64 atomics
//   variant 3: variant 2 + 4 replicas of the global histogram (cluster id & 3), summed by the last CTA
//   variant 4: variant 0 but the tail reads with __ldcg instead of volatile
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hist_tail_probe hist_tail_probe.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
namespace cg = cooperative_groups;
constexpr int kThreads = 256, kBins = 4096, kPer = kBins / kThreads;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

struct Stamps { unsigned long long t_first_entry, t_last_flush_issued, t_tail_begin, t_tail_read, t_tail_end, t_last_fence_done; };

__device__ __forceinline__ uint32_t rnd(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int VARIANT>
__global__ void __launch_bounds__(kThreads, 4)
k_probe(unsigned long long* hist, uint32_t* hist32, unsigned int* ticket, Stamps* st, int keys_per_cta, int lo_bin, int n_bins, int spin_us,
        unsigned long long* out_total) {
    __shared__ uint32_t s_hist[kBins];
    __shared__ unsigned int s_ticket;
    __shared__ unsigned long long s_warp[9];
    const int tid = threadIdx.x;
    const unsigned long long t0 = gtime();
    if (tid == 0) atomicMin(&st->t_first_entry, t0);
    for (int b = tid; b < kBins; b += kThreads) s_hist[b] = 0;
    __syncthreads();
    uint32_t seed = blockIdx.x * 7919u + tid * 104729u + 1u;
    for (int i = tid; i < keys_per_cta; i += kThreads) atomicAdd(&s_hist[lo_bin + rnd(seed) % n_bins], 1u);
    // all CTAs reach the flush at the same time, like the balanced sweep
    while (gtime() - st->t_first_entry < (unsigned long long)spin_us * 1000ull) { }
    __syncthreads();
    if (VARIANT == 2 || VARIANT == 3) {
        cg::cluster_group cl = cg::this_cluster();
        cl.sync();
        const unsigned r = cl.block_rank(), nr = cl.num_blocks();
        const int share = kBins / nr;                                   // 512 bins per CTA at 8
        unsigned long long* dst = hist + (VARIANT == 3 ? (size_t)((blockIdx.x / nr) & 3) * (kBins + 8) : 0);
        for (int b = r * share + tid; b < (int)(r + 1) * share; b += kThreads) {
            uint32_t v = 0;
            for (unsigned q = 0; q < nr; ++q) v += *cl.map_shared_rank(&s_hist[b], q);
            if (v) atomicAdd(dst + b, (unsigned long long)v);
        }
        cl.sync();
    } else if (VARIANT == 5 || VARIANT == 6) {
        unsigned long long* dst = hist + (size_t)(blockIdx.x & 3) * (kBins + 8);
        for (int b = tid; b < kBins; b += kThreads) { const uint32_t v = s_hist[b]; if (v) atomicAdd(dst + b, (unsigned long long)v); }
    } else if (VARIANT == 1) {
        for (int b = tid; b < kBins; b += kThreads) { const uint32_t v = s_hist[b]; if (v) atomicAdd(hist32 + b, v); }
    } else {
        for (int b = tid; b < kBins; b += kThreads) { const uint32_t v = s_hist[b]; if (v) atomicAdd(hist + b, (unsigned long long)v); }
    }
    if (tid == 0) atomicMax(&st->t_last_flush_issued, gtime());
    __threadfence();
    if (tid == 0) atomicMax(&st->t_last_fence_done, gtime());
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    if (tid == 0) st->t_tail_begin = gtime();
    unsigned long long local[kPer];
    if (VARIANT == 5 || VARIANT == 6) {
        uint32_t sum[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            unsigned long long v = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) v += __ldcg(hist + (size_t)q * (kBins + 8) + i * kThreads + tid);
            sum[i] = (uint32_t)v;
        }
        if (VARIANT == 6 && tid == 0) st->t_tail_read = gtime();
#pragma unroll
        for (int i = 0; i < kPer; ++i) s_hist[i * kThreads + tid] = sum[i];
        for (int q = 1; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < kPer; ++i) hist[(size_t)q * (kBins + 8) + i * kThreads + tid] = 0ull;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kPer; ++i) local[i] = s_hist[tid * kPer + i];
    } else if (VARIANT == 1) {
#pragma unroll
        for (int i = 0; i < kPer; ++i) local[i] = ((volatile uint32_t*)hist32)[tid * kPer + i];
    } else if (VARIANT == 3) {
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            unsigned long long s = 0;
            for (int q = 0; q < 4; ++q) s += __ldcg(hist + (size_t)q * (kBins + 8) + tid * kPer + i);
            local[i] = s;
        }
    } else if (VARIANT == 4) {
#pragma unroll
        for (int i = 0; i < kPer; ++i) local[i] = __ldcg(hist + tid * kPer + i);
    } else {
#pragma unroll
        for (int i = 0; i < kPer; ++i) local[i] = ((volatile unsigned long long*)hist)[tid * kPer + i];
    }
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) sum += local[i];
    if (tid == 0 && VARIANT != 6) st->t_tail_read = gtime();
    unsigned long long incl = sum;
    for (int o = 1; o < 32; o <<= 1) { unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((tid & 31) >= o) incl += v; }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    if (tid == 0) { unsigned long long t = 0; for (int w = 0; w < kThreads / 32; ++w) t += s_warp[w]; *out_total = t; *ticket = 0u; }
    const int reps = VARIANT == 3 ? 4 : 1;
    if (VARIANT == 5 || VARIANT == 6) {
#pragma unroll
        for (int i = 0; i < kPer; ++i) hist[i * kThreads + tid] = 0ull;
    } else
    for (int q = 0; q < reps; ++q)
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            if (VARIANT == 1) hist32[tid * kPer + i] = 0u; else hist[(size_t)q * (kBins + 8) + tid * kPer + i] = 0ull;
        }
    __syncthreads();
    if (tid == 0) st->t_tail_end = gtime();
}

template <int V>
static void run(const char* name, int grid, int keys, int lo, int nb, int spin, bool cluster) {
    unsigned long long* hist; uint32_t* hist32; unsigned int* ticket; Stamps* st; unsigned long long* total;
    cudaMalloc(&hist, 4 * (kBins + 8) * 8); cudaMalloc(&hist32, kBins * 4); cudaMalloc(&ticket, 4); cudaMalloc(&st, sizeof(Stamps)); cudaMalloc(&total, 8);
    cudaMemset(hist, 0, 4 * (kBins + 8) * 8); cudaMemset(hist32, 0, kBins * 4); cudaMemset(ticket, 0, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double acc[6] = {0, 0, 0, 0, 0, 0}; float ms_acc = 0; int n = 0;
    unsigned long long tot = 0;
    for (int it = 0; it < 12; ++it) {
        Stamps h = {~0ull, 0, 0, 0, 0, 0};
        cudaMemcpy(st, &h, sizeof(h), cudaMemcpyHostToDevice);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.stream = 0;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = cluster ? 8 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaEventRecord(e0);
        cudaError_t err = cudaLaunchKernelEx(&cfg, k_probe<V>, hist, hist32, ticket, st, keys, lo, nb, spin, total);
        cudaEventRecord(e1);
        if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&h, st, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemcpy(&tot, total, 8, cudaMemcpyDeviceToHost);
        if (it < 2) continue;
        const double b = (double)h.t_first_entry;
        acc[0] += (h.t_last_flush_issued - b) / 1e3; acc[1] += (h.t_last_fence_done - b) / 1e3; acc[2] += (h.t_tail_begin - b) / 1e3;
        acc[3] += (h.t_tail_read - b) / 1e3; acc[4] += (h.t_tail_end - b) / 1e3; ms_acc += ms; ++n;
    }
    printf("%-44s grid %4d keys/cta %5d bins [%d,+%d) spin %d us | flush issued %6.2f  fences done %6.2f  tail begin %6.2f  read %6.2f  end %6.2f us | kernel %6.2f us | total %llu (expect %llu)\n",
           name, grid, keys, lo, nb, spin, acc[0] / n, acc[1] / n, acc[2] / n, acc[3] / n, acc[4] / n, 1e3 * ms_acc / n, tot, (unsigned long long)grid * keys);
    cudaFree(hist); cudaFree(hist32); cudaFree(ticket); cudaFree(st); cudaFree(total);
}

int main() {
    // bracket sweep of ResNet-50: 592 CTAs, ~860 candidates per CTA over ~1000 fine bins
    run<0>("sweep  u64 atomics (current)", 592, 860, 1000, 1000, 20, false);
    run<5>("sweep  4 copies, coalesced tail", 592, 860, 1000, 1000, 20, false);
    run<6>("sweep  4 copies, coalesced (stamp after loads)", 592, 860, 1000, 1000, 20, false);
    run<5>("sample 4 copies, coalesced tail", 98, 4096, 1800, 300, 5, false);
    run<4>("sweep  u64 atomics, ldcg tail", 592, 860, 1000, 1000, 20, false);
    run<1>("sweep  u32 atomics", 592, 860, 1000, 1000, 20, false);
    run<2>("sweep  cluster-8 DSMEM reduce + u64", 592, 860, 1000, 1000, 20, true);
    run<3>("sweep  cluster-8 + 4 replicas", 592, 860, 1000, 1000, 20, true);
    run<0>("sweep  u64, all 4096 bins", 592, 4096, 0, 4096, 20, false);
    run<2>("sweep  cluster-8, all 4096 bins", 592, 4096, 0, 4096, 20, true);
    // sample kernel: 98 CTAs, 4096 sampled keys per CTA over ~300 coarse buckets
    run<0>("sample u64 atomics (current)", 98, 4096, 1800, 300, 5, false);
    run<1>("sample u32 atomics", 98, 4096, 1800, 300, 5, false);
    run<2>("sample cluster-8 (grid 104)", 104, 4096, 1800, 300, 5, true);
    run<0>("sample u64, no spin", 98, 4096, 1800, 300, 0, false);
    return 0;
}
