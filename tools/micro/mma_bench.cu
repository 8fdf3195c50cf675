// mma_bench.cu -- micro-benchmark: cycles per tcgen05.mma (M = 128, cta_group::1) as a function of N, operand kind,
// where A comes from (shared memory / tensor memory) and how many INDEPENDENT accumulators the instruction stream cycles
// through.  Answers: is a chain of MMAs into one accumulator latency-bound?   nvcc -arch=sm_100a -o mma_bench mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <int KIND>  // 0 = tf32, 1 = bf16
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred;
}

template <int KIND, bool TS, int VAR, int NACC>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, long long *out) {
    extern __shared__ __align__(1024) char smem_raw[];
    char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < (16384 + 32768) / 16; i += 128) reinterpret_cast<float4 *>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_slot;
    if (VAR == 0 ? tid == 0 : tid < 32) {
        const uint32_t fmt = KIND == 0 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + 16384));
        const uint32_t a_tmem = tb + 448;  // 64 columns at the top of tensor memory
        const uint32_t leader = VAR == 2 ? elect_one() : 1u;
        const long long t0 = clock64();
        // straight-line body: 8 MMAs per trip, every operand a loop-invariant register (+ a compile-time offset)
        const uint32_t d0 = tb, d1 = tb + (NACC > 1 ? N : 0), d2 = tb + (NACC > 2 ? 2 * N : 0), d3 = tb + (NACC > 2 ? 3 * N : (NACC > 1 ? N : 0));
        for (int it = 0; it < iters / 2; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t d = (j & 3) == 0 ? d0 : ((j & 3) == 1 ? d1 : ((j & 3) == 2 ? d2 : d3));
                bool go = true;
                if (VAR == 1) go = elect_one();
                if (VAR == 2) go = leader;
                if (go) {
                    if (TS) mma_ts<KIND>(d, a_tmem + (j & 3) * 8, db + (uint64_t)((j & 3) * 2), idesc, 1u);
                    else mma_ss<KIND>(d, da + (uint64_t)((j & 3) * 2), db + (uint64_t)((j & 3) * 2), idesc, 1u);
                }
            }
        }
        if (VAR == 0 || elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0 && tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int KIND, bool TS, int VAR, int NACC>
static void run(const char *name, int N, int grid) {
    long long *d_out, h[2];
    cudaMalloc(&d_out, 16);
    const int iters = 512;
    cudaFuncSetAttribute(bench<KIND, TS, VAR, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    bench<KIND, TS, VAR, NACC><<<grid, 128, 52 * 1024>>>(N, iters, d_out);
    bench<KIND, TS, VAR, NACC><<<grid, 128, 52 * 1024>>>(N, iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    printf("%-5s A=%s issue-variant=%d N=%3d accumulators=%d grid=%3d: issue %6.1f cyc/mma, complete %6.1f cyc/mma  (floor %d)  %s\n", name, TS ? "tmem" : "smem",
           VAR, N, NACC, grid, (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d_out);
}

int main() {
    for (int N : {32, 64, 128, 256}) {
        run<0, false, 0, 1>("tf32", N, 148);
        run<0, true, 0, 1>("tf32", N, 148);
        if (N <= 128) run<0, true, 0, 2>("tf32", N, 148);
        if (N <= 64) run<0, true, 0, 4>("tf32", N, 148);
        run<0, true, 2, 1>("tf32", N, 148);
        run<1, false, 0, 1>("bf16", N, 148);
    }
    return 0;
}
