// Version, error string, parameter layout queries, FFMA peak probe.
#include <stdarg.h>

#include "../../include/openobj_b200.h"
#include "oo_common.cuh"
#include "oo_layout.h"

namespace oo {
std::string& last_error() {
    static thread_local std::string e;
    return e;
}
int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}
}  // namespace oo

static_assert(OO_PSTRIDE == oo::PSTRIDE && OO_PCOUNT == oo::PCOUNT && OO_N_TENSORS == oo::NT, "header / layout mismatch");
static_assert(OO_TILE_RAYS == oo::RT && OO_NSAMP == oo::S && OO_CLIP == oo::C && OO_HIDDEN == oo::H, "header / layout mismatch");

extern "C" int oo_version(void) { return OO_ABI_VERSION; }
extern "C" const char* oo_last_error(void) { return oo::last_error().c_str(); }
extern "C" int oo_param_offset(int i) { return (i >= 0 && i < oo::NT) ? oo::kOff[i] : -1; }
extern "C" int oo_param_size(int i) { return (i >= 0 && i < oo::NT) ? oo::kSize[i] : -1; }

// ---- FP32 FMA-pipe peak probe: 8 independent FFMA chains per thread, 1024 threads per SM ----------
__global__ void __launch_bounds__(256) k_fma_peak(int iters, float* sink, float b, float c) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f,
          a6 = a0 + 6.f, a7 = a0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    const float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456f) sink[0] = s;
}
// launches n_sm*4 blocks of 256 threads; FLOPs = n_sm*4*256 * iters * 16 * 8 * 2
extern "C" int oo_fma_peak(int n_sm, int iters, float* sink, void* stream) {
    k_fma_peak<<<n_sm * 4, 256, 0, (cudaStream_t)stream>>>(iters, sink, 0.999f, 1e-3f);
    OO_LAUNCH_CHECK();
    return 0;
}
