// lost_common.cuh — image records shared by the LOST kernels (lost.cu, lost_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace b200p {

struct LostImageDev {
    long long feat_off, a_off, out_off;
    int n, dim0, dim1, img_h, img_w;
    float s0, s1;
    int tile_base;       // first CTA of this image in the Gram grid
    int tiles;           // 128-wide tiles per side
    int row_base;        // first row of this image in the stacked hi/lo operand arrays (direct mode: in the caller's array)
    int pair_base;       // first tile of this image in the symmetric (ti <= tj) tile list of the tensor-core Gram
    int pair2_base;      // the same for the 256x256 tiles of the CTA-pair kernel
    int pad_;
};

__device__ __forceinline__ unsigned long long lost_globaltimer() {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

__device__ __forceinline__ int find_image(const LostImageDev* __restrict__ meta, int n_images, int cta) {
    int lo = 0, hi = n_images - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (meta[mid].tile_base <= cta) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// tensor-core Gram (lost_tc.cu)
enum { LOST_TC_SINGLE = 0, LOST_TC_PAIR = 1, LOST_TC_PAIR_DIRECT = 2 };
size_t lost_tc_workspace_bytes(int n_images, long long total_patches, int d);
bool lost_tc_direct_ok(const float* d_feats, long long row_stride, int d, const b200p_lost_image_t* h_meta, int n_images);
struct Tile2;
struct LostGramPlan {
    alignas(64) unsigned char tm_hi[128], tm_lo[128];     // CUtensorMap storage (cuda.h stays out of this header)
    const Tile2* tab; const LostImageDev* d_meta;
    unsigned int* d_done;                                  // count-only: per-image completion counters (epilogue warps x tiles)
    int n_tiles2, n_tiles1, n_images, mode, d_pad, sms, nseg;
};
int lost_gram_prepare(LostGramPlan* gp, const float* d_feats, long long row_stride, int d, const LostImageDev* d_meta,
                      const std::vector<LostImageDev>& meta, long long total_patches, int n_max, void* ws, size_t ws_bytes,
                      int vec_ok, cudaStream_t st, int mode, bool count_only);
int lost_conv_warps();
int lost_gram_run(const LostGramPlan& gp, int t_begin, int t_end, float* A_base, int* d_degree, cudaStream_t st);

}  // namespace b200p
