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

// bf16 storage mode (pre-normalisation activations kept in bf16): round-to-nearest-even packing of two floats into one 32-bit
// word (low half = first value) and the exact widening back
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// 4 consecutive channels of a row that is stored either as fp32 or as bf16 (`ld` and `c` in elements)
__device__ __forceinline__ float4 load4_any(const void *base, int is_bf16, size_t row, int ld, int c) {
    if (is_bf16) {
        const uint2 w = *reinterpret_cast<const uint2 *>(reinterpret_cast<const unsigned short *>(base) + row * ld + c);
        return make_float4(bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y));
    }
    return *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(base) + row * ld + c);
}

// 2^x and 1/x on the SFU, one instruction each (arguments of the softmax are <= 0, results in (0, 1]; the denominators >= 1)
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// out[i] (+)= sum_c part[c][i], double accumulation in an order that depends only on (chunks, E): thread (e, g) of a CTA owns
// element blockIdx.x*E + e and the chunks g, g+G, ... (G = 256/E) with four independent loads in flight; the G group sums
// are combined by a shared-memory tree.  Loads are coalesced over e (E*4 bytes per group and chunk).
template <int E>
__global__ void __launch_bounds__(256) reduce_parts_kernel(const float *__restrict__ part, int chunks, long long n,
                                                           float *__restrict__ out, int accumulate) {
    constexpr int G = 256 / E;
    __shared__ double red[G][E];
    const int e = threadIdx.x % E, g = threadIdx.x / E;
    const long long i = (long long)blockIdx.x * E + e;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (i < n) {
        int c = g;
        for (; c + 3 * G < chunks; c += 4 * G) {
            const float a = part[(size_t)c * n + i], b = part[(size_t)(c + G) * n + i], d = part[(size_t)(c + 2 * G) * n + i],
                        f = part[(size_t)(c + 3 * G) * n + i];
            s0 += (double)a; s1 += (double)b; s2 += (double)d; s3 += (double)f;
        }
        for (; c < chunks; c += G) s0 += (double)part[(size_t)c * n + i];
    }
    red[g][e] = (s0 + s1) + (s2 + s3);
    __syncthreads();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        if (g < o) red[g][e] += red[g + o][e];
        __syncthreads();
    }
    if (g == 0 && i < n) out[i] = accumulate ? out[i] + (float)red[0][e] : (float)red[0][e];
}
// launch helper: narrow CTAs (8 elements x 32 chunk groups) for small outputs, wide ones (32 x 8) otherwise
// (few chunks: wide CTAs of 128 x 2 -- with 3 partials of a 1024 x 1024 weight gradient five of the eight chunk groups of
// the 32 x 8 shape had nothing to do, 60 us per launch)
inline void launch_reduce_parts(const float *part, int chunks, long long n, float *out, int accumulate, cudaStream_t st) {
    if (chunks <= 8 && n > 8192) reduce_parts_kernel<128><<<ceil_div(n, 128), 256, 0, st>>>(part, chunks, n, out, accumulate);
    else if (n <= 8192) reduce_parts_kernel<8><<<ceil_div(n, 8), 256, 0, st>>>(part, chunks, n, out, accumulate);
    else reduce_parts_kernel<32><<<ceil_div(n, 32), 256, 0, st>>>(part, chunks, n, out, accumulate);
}

}  // namespace pu
