// Does a programmatic dependent kernel get co-scheduled beside a persistent primary that fills the SMs' shared memory?
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <bool CLUSTER>
__global__ void __launch_bounds__(448, 1) primary(unsigned long long* t, int spin_us, int trigger) {
    extern __shared__ unsigned char smem[];
    if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const unsigned long long t0 = gtime();
    if (threadIdx.x == 0) smem[0] = 1;
    while (gtime() - t0 < (unsigned long long)spin_us * 1000ull) { }
    if (threadIdx.x == 0 && blockIdx.x == 0) { t[0] = t0; t[1] = gtime(); }
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(448, 1) primary_cluster(unsigned long long* t, int spin_us, int trigger) {
    extern __shared__ unsigned char smem[];
    if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const unsigned long long t0 = gtime();
    if (threadIdx.x == 0) smem[0] = 1;
    while (gtime() - t0 < (unsigned long long)spin_us * 1000ull) { }
    if (threadIdx.x == 0 && blockIdx.x == 0) { t[0] = t0; t[1] = gtime(); }
}
__global__ void secondary(unsigned long long* t) {
    extern __shared__ unsigned char smem[];
    if (threadIdx.x == 0) { smem[0] = 1; atomicMin(&t[2], gtime()); atomicMax(&t[3], gtime()); }
}
int main(int argc, char** argv) {
    const int prim_smem = argc > 1 ? atoi(argv[1]) : 197888, sec_smem = argc > 2 ? atoi(argv[2]) : 21500, sec_threads = argc > 3 ? atoi(argv[3]) : 384;
    const int carve = argc > 4 ? atoi(argv[4]) : 100;
    unsigned long long* d; cudaMalloc(&d, 32);
    cudaFuncSetAttribute(primary<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, prim_smem);
    cudaFuncSetAttribute(primary_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, prim_smem);
    cudaFuncSetAttribute(secondary, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (carve >= 0) {
        cudaFuncSetAttribute(primary<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(primary_cluster, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(secondary, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
    for (int cluster = 0; cluster < 2; ++cluster)
    for (int trigger = 0; trigger < 2; ++trigger)
    for (int pdl = 0; pdl < 2; ++pdl) {
        unsigned long long h[4] = {0, 0, ~0ull, 0};
        cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice);
        if (cluster) primary_cluster<<<148, 448, prim_smem>>>(d, 300, trigger);
        else primary<false><<<148, 448, prim_smem>>>(d, 300, trigger);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(256); cfg.blockDim = dim3(sec_threads); cfg.dynamicSmemBytes = sec_smem; cfg.stream = 0;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, secondary, d);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("cluster=%d trigger=%d pdl=%d (%s): primary ran %.1f us; first secondary CTA started %.1f us after the primary's start, last %.1f us\n",
               cluster, trigger, pdl, cudaGetErrorString(e), (h[1] - h[0]) / 1e3, ((double)h[2] - (double)h[0]) / 1e3, ((double)h[3] - (double)h[0]) / 1e3);
    }
    return 0;
}
