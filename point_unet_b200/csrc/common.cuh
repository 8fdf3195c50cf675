// common.cuh -- shared helpers for libpointunet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pointunet_b200.h"

#ifndef __CUDA_ARCH_LIST__
#endif

namespace pu {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of this
constexpr int kWarp = 32;

extern thread_local int g_last_cuda_error;
extern unsigned long long g_launch_count;

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return PU_ERR_CUDA;
}

inline void count_launch(int n = 1) { __atomic_fetch_add(&g_launch_count, (unsigned long long)n, __ATOMIC_RELAXED); }

#define PU_CUDA_TRY(expr)                                  \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) return ::pu::cuda_fail(_e); \
    } while (0)

// checks the launch just issued (configuration errors surface here; execution errors at the next sync)
#define PU_LAUNCH_CHECK()                                  \
    do {                                                   \
        ::pu::count_launch();                              \
        cudaError_t _e = cudaGetLastError();               \
        if (_e != cudaSuccess) return ::pu::cuda_fail(_e); \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// 128-bit streaming accessors: data touched once bypasses L1 (guide: Guideline 13/14)
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

}  // namespace pu
