// Error plumbing shared by all translation units of libopenobj_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include <string>

namespace oo {
std::string& last_error();
int fail(int code, const char* fmt, ...);

// Function attributes (cudaFuncSetAttribute) and SM counts belong to a device, not to the process: one value per device
// ordinal, addressed by the calling thread's current device.
struct PerDevice {
    size_t v[64] = {};
    size_t& cur() {
        int d = 0;
        cudaGetDevice(&d);
        return v[d & 63];
    }
};
}  // namespace oo

#define OO_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return oo::fail(-100 - (int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                               \
    } while (0)

#define OO_LAUNCH_CHECK() OO_CUDA(cudaGetLastError())

#define OO_REQUIRE(cond, ...)                    \
    do {                                         \
        if (!(cond)) return oo::fail(-1, __VA_ARGS__); \
    } while (0)
