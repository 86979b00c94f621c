// Same-address / same-line atomic throughput at the end of a balanced sweep: what does the histogram flush cost?
// Every CTA adds `nz` non-zero bins (random subset of n_bins) into replica (blockIdx % R) of a global histogram whose bin b
// lives at word b * stride (u64 or u32 words), plus `single` atomics on one shared counter; stamps: last "flush issued",
// last "fence done".  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hist_flush_probe hist_flush_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
constexpr int kThreads = 256;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
struct Stamps { unsigned long long t0, t_issue, t_fence; };
template <typename T>
__global__ void __launch_bounds__(kThreads, 4)
k_flush(T* hist, unsigned long long* counter, Stamps* st, int n_bins, int nz_per_thread, int stride, int R, size_t rep_words, int single, int spin_us) {
    const int tid = threadIdx.x;
    const unsigned long long t0 = gtime();
    if (tid == 0) atomicMin(&st->t0, t0);
    __syncthreads();
    while (gtime() - st->t0 < (unsigned long long)spin_us * 1000ull) { }
    __syncthreads();
    T* dst = hist + (size_t)(blockIdx.x % R) * rep_words;
    uint32_t s = blockIdx.x * 7919u + tid * 104729u + 1u;
    for (int i = 0; i < nz_per_thread; ++i) {
        s = s * 1664525u + 1013904223u;
        const int b = (tid * nz_per_thread + i + (s >> 8) % 3) % n_bins;          // ~distinct bins per CTA
        atomicAdd(dst + (size_t)b * stride, (T)1);
    }
    if (single && tid == 0) atomicAdd(counter, 1ull);
    if (tid == 0) atomicMax(&st->t_issue, gtime());
    __threadfence();
    if (tid == 0) atomicMax(&st->t_fence, gtime());
}
template <typename T>
static void run(const char* name, int grid, int n_bins, int nzpt, int stride, int R, int single) {
    const size_t rep_words = (size_t)n_bins * stride + 64;
    T* hist; unsigned long long* counter; Stamps* st;
    cudaMalloc(&hist, rep_words * R * sizeof(T)); cudaMalloc(&counter, 8); cudaMalloc(&st, sizeof(Stamps));
    cudaMemset(hist, 0, rep_words * R * sizeof(T)); cudaMemset(counter, 0, 8);
    double a = 0, f = 0; int n = 0;
    for (int it = 0; it < 12; ++it) {
        Stamps h = {~0ull, 0, 0};
        cudaMemcpy(st, &h, sizeof(h), cudaMemcpyHostToDevice);
        k_flush<T><<<grid, kThreads>>>(hist, counter, st, n_bins, nzpt, stride, R, rep_words, single, 20);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s failed\n", name); return; }
        cudaMemcpy(&h, st, sizeof(h), cudaMemcpyDeviceToHost);
        if (it < 2) continue;
        a += (h.t_issue - h.t0) / 1e3 - 20.0; f += (h.t_fence - h.t0) / 1e3 - 20.0; ++n;
    }
    printf("%-40s grid %4d bins %5d nz/cta %5d stride %2d R %2d single %d | issued +%6.2f us  fenced +%6.2f us\n", name, grid, n_bins, nzpt * kThreads, stride, R, single, a / n, f / n);
    cudaFree(hist); cudaFree(counter); cudaFree(st);
}
int main() {
    run<unsigned long long>("nothing (spin exit skew)", 592, 1024, 0, 1, 1, 0);
    run<unsigned long long>("one shared counter only", 592, 1024, 0, 1, 1, 1);
    run<unsigned long long>("u64 1024 bins", 592, 1024, 3, 1, 1, 0);
    run<unsigned long long>("u64 1024 bins + counter", 592, 1024, 3, 1, 1, 1);
    run<uint32_t>("u32 1024 bins", 592, 1024, 3, 1, 1, 0);
    run<unsigned long long>("u64 1024 bins, one per 32B sector", 592, 1024, 3, 4, 1, 0);
    run<unsigned long long>("u64 1024 bins, one per 128B line", 592, 1024, 3, 16, 1, 0);
    run<unsigned long long>("u64 1024 bins R=2", 592, 1024, 3, 1, 2, 0);
    run<unsigned long long>("u64 1024 bins R=4", 592, 1024, 3, 1, 4, 0);
    run<unsigned long long>("u64 1024 bins R=8", 592, 1024, 3, 1, 8, 0);
    run<uint32_t>("u32 1024 bins R=8", 592, 1024, 3, 1, 8, 0);
    run<uint32_t>("u32 1024 bins R=16", 592, 1024, 3, 1, 16, 0);
    run<unsigned long long>("u64 4096 bins nz 4096", 592, 4096, 16, 1, 1, 0);
    run<unsigned long long>("u64 4096 bins nz 4096 R=8", 592, 4096, 16, 1, 8, 0);
    run<unsigned long long>("u64 256 bins nz 256", 592, 256, 1, 1, 1, 0);
    run<unsigned long long>("sample: 98 CTAs 300 bins", 98, 300, 1, 1, 1, 0);
    run<unsigned long long>("sample: 98 CTAs 300 bins R=8", 98, 300, 1, 1, 8, 0);
    return 0;
}
