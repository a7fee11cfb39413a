// cgl_internal.cuh -- helpers shared by the translation units of libcgl_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cgl_b200.h"
#include "cgl_bits.cuh"

namespace cgl {

void set_error(const char *fmt, ...);
int sm_count();

#define CGL_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            cgl::set_error(__VA_ARGS__);        \
            return (code);                      \
        }                                       \
    } while (0)

#define CGL_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            cgl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                        \
            return (int)_e;                                                                  \
        }                                                                                    \
    } while (0)

#define CGL_LAUNCH_CHECK()                                                                  \
    do {                                                                                    \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) {                                                            \
            cgl::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),      \
                           __FILE__, __LINE__);                                             \
            return (int)_e;                                                                 \
        }                                                                                   \
    } while (0)

// Kernel attributes (cudaFuncSetAttribute) belong to the device that is current when they are set: a
// process that drives several GPUs has to configure each of them once.
struct PerDeviceOnce {
    bool done[64] = {};
    bool first()
    {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

static inline cudaStream_t as_stream(cgl_stream_t s) { return (cudaStream_t)s; }

// grid size for a grid-stride kernel over `n` items with `threads` per block:
// enough blocks to cover n once, capped at a multiple of the SM count (148 on B200).
static inline unsigned grid_for(uint64_t n, unsigned threads, unsigned blocks_per_sm = 16)
{
    uint64_t need = (n + threads - 1) / threads;
    uint64_t cap = (uint64_t)sm_count() * blocks_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t ld_nc_u32(const uint32_t *p) { return __ldg(p); }
#endif

}  // namespace cgl
