// lost.cu — placeholder until the LOST kernels land (K6/K7).
#include "common.cuh"
using namespace b200p;
extern "C" int b200p_lost_workspace_bytes(int, int64_t, int64_t, int64_t* out) { if (out) *out = 0; return B200P_OK; }
extern "C" int b200p_lost_batched(int, const float*, int64_t, int, const b200p_lost_image_t*, int, int, float*, int32_t*,
                                  int32_t*, float*, int32_t*, void*, int64_t, int, void*) {
    set_error("lost_batched: not implemented yet"); return B200P_ESTATE;
}
