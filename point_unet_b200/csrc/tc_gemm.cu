// tc_gemm.cu -- 5th-generation tensor-core (tcgen05 + TMEM) path for the wide contractions of PointSegment.
//
//   pu_tc_linear_fwd        y[M,N] (+)= x[M,K] wt[N,K]^T + bias   (1x1 conv / dense / dgrad at >= 32 channels)
//   pu_tc_att_pooling_fwd/_bwd   fused FC + softmax over K + weighted sum and its gradient
//   pu_tc_wgrad             dW = x^T dy
//
// Precision.  The reference computes these GEMMs in fp32 and the parity bar is 1e-3 relative, so the default mode is
// "3xTF32": every fp32 operand is split as a = hi + lo with hi = round-to-tf32(a), lo = a - hi (exact), and the
// tensor core accumulates hi*hi + hi*lo + lo*hi in fp32 (TMEM).  The dropped lo*lo term is <= 2^-22 relative, i.e.
// fp32-class accuracy at one third of the tf32 tensor rate -- still several times the CUDA-core fp32 peak.
// mode 1 = plain TF32 (one MMA per k-step), for the stated reduced-precision tolerance.
//
// Operand layout shared by all kernels: K-major SWIZZLE_128B tiles (rows of 32 fp32 = 128 B; 16-byte chunk c of row r is
// stored at chunk c ^ (r % 8) inside its 1024-byte 8-row atom) -- the canonical UMMA layout (cute::UMMA::Layout_K_SW128_Atom).
// The hi/lo split needs a pass through registers anyway, so the tiles are written with st.shared (coalesced loads,
// conflict-free swizzled stores) and published to the async proxy with fence.proxy.async; accumulators live in TMEM and
// come back through tcgen05.ld (32 lanes x 16 columns per warp instruction).  Kernel anatomy: see tc_persist_kernel.
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)
#include <float.h>
#include <stdlib.h>

#include "common.cuh"

namespace pu {
namespace tc {

constexpr int BM = 128;  // UMMA_M (cta_group::1)
constexpr int BK = 32;   // fp32 elements per 128-byte swizzle row
constexpr int UMMA_K = 8;  // tf32: 32 bytes per instruction

struct Params {
    const float *A; int lda;    // [M,K]
    const float *Bt; int ldb;   // [N,K]  (K-major, i.e. the transposed weight)
    float *C; int ldc;          // [M,N]
    const float *bias;
    long long M; int N, K;
    int accumulate;
    float *stat_sum, *stat_m2;  // [row tiles, N] or null
    int mode;                   // 3 = 3xTF32 (fp32-class), 1 = TF32
    int *error_flag;            // set to 1 if an mbarrier wait timed out (never hangs the GPU)
    int c_bf16;                 // EPI_STORE: C is stored as bf16 (ldc in elements; never with accumulate)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// bounded wait: returns false on timeout.  PU_MBAR_HINT=1 builds the variant with a suspend-time hint (the hardware parks the
// warp until the phase completes or ~10 ms pass instead of returning after the short default limit).  Measured on B200
// (round 2): no gain for the weight-gradient kernel and 5 % LOSS for the persistent linear kernel (slower wake-up), so the
// plain polling loop stays the default.
#ifndef PU_MBAR_HINT
#define PU_MBAR_HINT 0
#endif
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#if PU_MBAR_HINT
    for (int it = 0; it < 64; ++it) {   // 64 x ~10 ms: a pipeline bug still ends in the error flag, not in a hang
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(0x989680u)
            : "memory");
        if (ok) return true;
    }
#else
    for (int it = 0; it < (1 << 22); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return true;
    }
#endif
    return false;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (= 1, unused for swizzled K-major) | [32,46) SBO >> 4 (8 rows * 128 B)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from tensor memory (lane = row, one 32-bit column per k element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
// one lane of a CONVERGED warp; the predicate is taken once and reused (`if (leader) tcgen05.mma ...`): ptxas then emits the
// MMAs back to back.  Issuing from inside a divergent `if (lane == 0)` region instead wraps every single MMA in an
// ELECT / BRA.U.ANY loop and costs ~45 cycles per instruction (measured, tools/micro/mma_bench.cu) -- more than a whole
// 128 x 64 x 8 MMA takes on the tensor core (32 cycles).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// round-to-nearest (ties away from zero) to 10 mantissa bits == cvt.rna.tf32.f32 for finite values and infinities, in two
// integer instructions (ptxas expands the cvt into four); a NaN stays a NaN through the lo = a - hi term
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_rn(float a) { return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u); }

// byte offset of 16-byte chunk c (0..7) of row r inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

// =============================================================================================================
// persistent, warp-specialised kernel.  grid = a multiple of the SM count; every CTA walks 128-row tiles.
//   warps 0-3  producers: global -> registers (prefetched one k-block ahead) -> hi/lo split -> swizzled smem ring;
//              thread 0 additionally issues the MMAs of the stage it just helped to fill (one elected issuer)
//   warps 4-7  epilogue: wait for the tile's accumulator (TMEM, double-buffered), tcgen05.ld, release the buffer,
//              then finish the tile (bias/accumulate/statistics, or the att_pooling softmax) while the tensor core
//              already works on the next tile
//   the weight operand (<= 96 KB as hi+lo tf32 pairs) is split once and stays resident in shared memory.
// Three epilogues share the pipeline:
//   EPI_STORE    y (+)= x wt^T + bias, optional per-tile batch-norm partials           (pu_tc_linear_fwd)
//   EPI_ATT_FWD  f_agg[p,c] = sum_k x[p,k,c] softmax_k(x w)[c]                          (pu_att_pooling_fwd, K = 16)
//   EPI_ATT_BWD  d_act = s (g x - sum_k g x s),  dx_direct = g s                        (pu_att_pooling_bwd)
//   EPI_ATT_BWD_F  the same plus the dgrad through the FC in the SAME kernel (d = 64): d_act goes back into tensor memory as
//                the A operand of a second MMA against the resident weight, dx = g s + d_act w^T leaves once -- the separate
//                accumulate-GEMM (read d_act, read dx_direct, write dx: 2.2 GB per launch at level 1) disappears
enum { EPI_STORE = 0, EPI_ATT_FWD = 1, EPI_ATT_BWD = 2, EPI_ATT_BWD_F = 3 };
constexpr int MAX_TA = 4;        // operand-A stages in TENSOR memory (64 columns each: hi 32 + lo 32)
constexpr int MAX_RAW = 6;       // raw fp32 ring of the weight-gradient kernel (cp.async, no registers held in flight)
constexpr int MAX_RAW_P = 10;    // raw fp32 ring of the persistent kernel: up to 9 k-blocks (144 KB) in flight per SM
constexpr int P_THREADS = 256;   // 8 producer warps
constexpr int E_THREADS = 512;   // 16 epilogue warps: the epilogue is a chain of dependent ALU work, it needs the thread-level parallelism

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid) {
    const uint32_t d = smem_u32(smem_dst);
    const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int pending) {  // wait until at most `pending` groups are in flight
    switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        case 7: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 8;" ::: "memory"); break;
    }
}

struct Params2 {
    int raw_depth;             // cp.async raw-ring depth D (k-blocks of 16 KB in flight per CTA = D - 1)
    const char *Bp;            // streamed variant: packed hi/lo swizzled weight image (tc_pack_weight_kernel)
    Params g;                  // the GEMM proper (A = x rows, Bt = weight in K-major form, C, bias, stats, mode)
    const float *X; int ldx;   // att: feature_set rows (same memory as A)
    const float *G; int ldg;   // att bwd: upstream gradient [M/16, N]
    float *OUT; int ldo;       // att fwd: f_agg [M/16, N];  att bwd: dx_direct [M, N] (fused: the complete dx)
    const float *W2; int ldw2; // EPI_ATT_BWD_F: the FC kernel in its own orientation w[c_in][j_out] (= K-major B operand of d_act w^T)
    long long ntiles;
#ifdef PU_TC_TIMELINE
    long long *timeline;       // development builds only (tools/tc_timeline.py): per-role clock64 stamps of CTA (0, 0)
#endif
};

// Per-role timeline of the persistent kernel, compiled in only with -DPU_TC_TIMELINE (the product build contains none of
// it).  CTA (0,0) stamps clock64() at the hand-off points of its first TL_ITEMS work items: role r, item i, event e ->
// timeline[(r * TL_ITEMS + i) * TL_EVENTS + e].  Roles: 0 loader, 1 converter (warp 0), 2 issuer, 3 / 4 epilogue groups.
#ifdef PU_TC_TIMELINE
constexpr int TL_ITEMS = 256, TL_EVENTS = 4, TL_ROLES = 5;
#define PU_TL(role, item, ev)                                                                              \
    do {                                                                                                   \
        if (q.timeline && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (item) < TL_ITEMS) \
            q.timeline[((size_t)(role) * TL_ITEMS + (item)) * TL_EVENTS + (ev)] = clock64();               \
    } while (0)
#else
#define PU_TL(role, item, ev) do { } while (0)
#endif

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival per WARP: every mbarrier.arrive is an atomic on one shared-memory word, and arrivals of different warps on the
// same barrier serialise -- with 256-512 arriving threads per 16 KB hand-off the barriers alone cost as many cycles as the
// hand-off has at full HBM speed (measured: the weight-gradient kernels sat at ~1100 cycles per 32-row step whatever the
// operand path).  __syncwarp orders the lanes' preceding shared / tensor-memory accesses before lane 0's releasing arrive.
// Barriers that use this are initialised with a count of (threads / 32).  PU_WARP_ARRIVE=0 restores per-thread arrivals.
#ifndef PU_WARP_ARRIVE
#define PU_WARP_ARRIVE 1
#endif
constexpr int ARRIVE_DIV = PU_WARP_ARRIVE ? 32 : 1;
__device__ __forceinline__ void mbar_arrive_warp(uint64_t *bar) {
#if PU_WARP_ARRIVE
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
#else
    mbar_arrive(bar);
#endif
}

// ---- streamed weight operand: pre-split / pre-swizzled image in global memory, fetched with 1-D bulk copies (TMA engine)
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// packed image: [n slab][k block][hi: BN rows x 128 B swizzled | lo: same]; one thread per 16-byte chunk
__global__ void __launch_bounds__(256) tc_pack_weight_kernel(const float *__restrict__ Bt, int ldb, int N, int K, int BN, int nkb,
                                                             int nslabs, int split, char *__restrict__ out) {
    const long long total = (long long)nslabs * nkb * BN * 8;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(t & 7);
        const int r = (int)((t >> 3) % BN);
        const long long sk = (t >> 3) / BN;  // slab * nkb + kb
        const int kb = (int)(sk % nkb), slab = (int)(sk / nkb);
        const int gn = slab * BN + r, gk = kb * BK + c * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gn < N && gk < K) v = *reinterpret_cast<const float4 *>(Bt + (size_t)gn * ldb + gk);
        char *base = out + (size_t)sk * (2 * BN * 128);
        const uint32_t off = sw128(r, c);
        if (split) {
            const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
            *reinterpret_cast<float4 *>(base + off) = h;
            *reinterpret_cast<float4 *>(base + BN * 128 + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        } else {
            *reinterpret_cast<float4 *>(base + off) = v;
            *reinterpret_cast<float4 *>(base + BN * 128 + off) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// 8 converter warps, 16 epilogue warps, 1 MMA-issuer warp, 1 TMA warp (x), 1 weight-fetch warp (streamed variant)
constexpr int PERSIST_THREADS = P_THREADS + E_THREADS + 96;

template <int BN, int EPI, bool STREAM>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) tc_persist_kernel(const Params2 q, const __grid_constant__ CUtensorMap tmap_a) {
    const Params &p = q.g;
    constexpr int A_BYTES = BM * 128;                 // one raw 128 x 32 fp32 k-block
    constexpr int B_KB = 2 * BN * 128;                // hi + lo of one k-block of the weight
    // operand-A stages in tensor memory = weight k-block stages of the streamed variant.  BN = 128: three (96 KB of weights
    // beside the staging tiles and a 3-deep raw ring); two left only ONE weight fetch in flight and every k-block waited a
    // full L2 round trip (~1.1 us against 0.39 us of MMA work).
    constexpr int TA = BN <= 64 ? MAX_TA : 3;
    constexpr int ACC_COLS = 2 * BN;                  // double-buffered accumulator; A stage s lives at column ACC_COLS + 64 s
    constexpr int TMEM_COLS = 512;                    // one CTA per SM: take all of tensor memory
    static_assert(ACC_COLS + TA * 64 <= TMEM_COLS, "tensor memory budget");
    constexpr bool FUSED = EPI == EPI_ATT_BWD_F;      // second MMA (d_act w^T) inside the epilogue
    constexpr int A2_COL = ACC_COLS + TA * 64;        // operand A of the second MMA: BN columns hi, BN columns lo
    static_assert(!FUSED || (!STREAM && BN == 64 && A2_COL + 2 * BN <= TMEM_COLS), "fused att_pooling backward: d = 64 only");
    constexpr int EC = BN > 64 ? 64 : BN;             // epilogue works on EC columns at a time (staging tile fits smem)
    constexpr int NPASS = BN / EC;
    constexpr int LDT = EC + 4;
    extern __shared__ __align__(1024) char smem_raw[];
    // 1024-byte alignment by OFFSET (not by an integer round trip): the pointer stays in the shared address space, so the
    // compiler emits LDS/STS with 32-bit addresses instead of generic LD/ST
    char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t stage_free[MAX_TA], stage_ready[MAX_TA], b_full[MAX_TA], acc_full[2], acc_empty[2];
    __shared__ uint64_t raw_full[MAX_RAW_P], raw_free[MAX_RAW_P];  // TMA landed a k-block / all converters have read it
    __shared__ uint64_t a2_free, acc2_full[2];                      // fused att backward: second-MMA operand / result hand-offs
    __shared__ uint32_t tmem_base_slot;
    __shared__ int s_err;
    __shared__ float s_red[2 * 16 * 64];  // statistics partials: [16 row lanes][EC][2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = (p.K + BK - 1) / BK;
    const int n0 = blockIdx.y * BN;
    const bool split = p.mode == 3;
    const int D = q.raw_depth;
    // att epilogues with d <= 64 take their x values from the raw ring (the tile's k-blocks stay resident until the epilogue
    // has used them) instead of re-reading them through L2: no load latency in the epilogue's critical path
    const bool x_ring = EPI != EPI_STORE && !STREAM && nkb <= 2 && p.N == p.K && D >= 7;
    constexpr bool ATT_BWD = EPI == EPI_ATT_BWD || EPI == EPI_ATT_BWD_F;
    char *raw_ring = smem;                            // D raw k-blocks, filled by TMA
    char *b_res = smem + (size_t)D * A_BYTES;         // resident: nkb k-blocks; streamed: TA k-blocks
    char *b2_res = b_res + (size_t)(STREAM ? TA : nkb) * B_KB;   // fused att backward: second resident weight image
    float *tile = reinterpret_cast<float *>(b2_res + (FUSED ? (size_t)nkb * B_KB : 0));  // epilogue staging [BM][LDT]
    const char *b_packed = STREAM ? q.Bp + (size_t)blockIdx.y * nkb * B_KB : nullptr;

    if (tid == 0) {
        for (int i = 0; i < MAX_TA; ++i) { mbar_init(&stage_free[i], 1); mbar_init(&stage_ready[i], P_THREADS / ARRIVE_DIV); mbar_init(&b_full[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], E_THREADS / 2 / ARRIVE_DIV); }
        for (int i = 0; i < MAX_RAW_P; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_free[i], (P_THREADS + (x_ring ? E_THREADS / 2 : 0)) / ARRIVE_DIV); }
        mbar_init(&a2_free, 1); mbar_init(&acc2_full[0], 1); mbar_init(&acc2_full[1], 1);
        s_err = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if constexpr (!STREAM) {
        // resident weight: all threads load + split every k-block once (fused att backward: both orientations)
        for (int kb = 0; kb < (FUSED ? 2 * nkb : nkb); ++kb) {
            const bool second = kb >= nkb;
            const int kbb = second ? kb - nkb : kb;
            const float *wsrc = second ? q.W2 : p.Bt;
            const int wld = second ? q.ldw2 : p.ldb;
            char *b_hi = (second ? b2_res : b_res) + (size_t)kbb * B_KB, *b_lo = b_hi + BN * 128;
            for (int idx = tid; idx < BN * 8; idx += PERSIST_THREADS) {
                const int r = idx >> 3, c = idx & 7;
                const int gn = n0 + r, gk = kbb * BK + c * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gn < p.N && gk < p.K) v = *reinterpret_cast<const float4 *>(wsrc + (size_t)gn * wld + gk);
                const uint32_t off = sw128(r, c);
                if (split) {
                    const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
                    *reinterpret_cast<float4 *>(b_hi + off) = h;
                    *reinterpret_cast<float4 *>(b_lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                } else {
                    *reinterpret_cast<float4 *>(b_hi + off) = v;
                }
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp < P_THREADS / 32) {
        // ======================= converters =======================
        // work items = (tile, k-block) pairs in order.
        //   global --TMA (one elected lane of the loader warp, 128B-swizzled box of 128 rows x 32 floats)--> raw ring in shared
        //   memory (D k-blocks, up to 160 KB in flight per SM, no registers, no per-thread copy instructions)
        //          --LDS--> registers: hi/lo tf32 split --tcgen05.st--> operand-A stage in TENSOR memory.
        // The x operand never exists as a hi/lo image in shared memory: the tensor core reads A from TMEM and only the
        // (small, resident) weight from shared memory.  The swizzle puts the 16-byte chunk c of row r at
        // r*128 + ((c ^ (r & 7)) << 4), so the row-per-lane reads below are bank-conflict free.
        // Mapping (tensor-memory lanes): warp w owns rows 32 (w & 3) .. +31 (its TMEM lane quarter), lane = row, and the
        // 16 columns 16 (w >> 2) .. +15 of the k-block.
        long long cur_tile = blockIdx.x;
        int cur_kb = 0, rslot = 0, ruse = 0, slot = 0, use = 0;
        bool ok = true;
        const int crow = (warp & 3) * 32 + lane, chalf = warp >> 2;
        uint32_t c_off[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) c_off[j] = (uint32_t)(crow * 128 + (((chalf * 4 + j) ^ (crow & 7)) << 4));
        const uint32_t t_a = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(ACC_COLS + chalf * 16);
#ifdef PU_TC_TIMELINE
        int tl_item = 0;
#endif
        while (cur_tile < q.ntiles) {
            ok = mbar_wait(&raw_full[rslot], (uint32_t)(ruse & 1)) && ok;                        // the k-block has landed
#ifdef PU_TC_TIMELINE
            if (warp == 0) PU_TL(1, tl_item, 0);
#endif
            if (use >= 1) ok = mbar_wait(&stage_free[slot], (uint32_t)((use - 1) & 1)) && ok;    // MMAs that read the stage retired
#ifdef PU_TC_TIMELINE
            if (warp == 0) PU_TL(1, tl_item, 1);
#endif
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const char *src = raw_ring + (size_t)rslot * A_BYTES;
            float v[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 t = *reinterpret_cast<const float4 *>(src + c_off[j]);
                v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
            }
            const uint32_t ta = t_a + (uint32_t)(slot * 64);
            if (split) {
                float h[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { h[j] = tf32_rn(v[j]); v[j] -= h[j]; }
                tmem_st16(ta, h);
                tmem_st16(ta + 32, v);
            } else {
                tmem_st16(ta, v);
            }
            // the tensor-memory stores consumed every loaded value, so the shared-memory reads have completed: only now may
            // the loader refill the slot (an arrive issued right behind the LDS can overtake it in the memory pipeline)
            mbar_arrive_warp(&raw_free[rslot]);
#ifdef PU_TC_TIMELINE
            if (warp == 0) PU_TL(1, tl_item, 2);
#endif
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive_warp(&stage_ready[slot]);  // hand the stage to the issuer; do not wait for it
#ifdef PU_TC_TIMELINE
            if (warp == 0) PU_TL(1, tl_item, 3);
            ++tl_item;
#endif
            if (++cur_kb == nkb) { cur_kb = 0; cur_tile += gridDim.x; }
            if (++rslot == D) { rslot = 0; ++ruse; }
            if (++slot == TA) { slot = 0; ++use; }
        }
        if (!ok) s_err = 1;
    } else if (warp == (P_THREADS + E_THREADS) / 32 + 1) {
        // ======================= loader: one elected lane feeds the raw ring with TMA tile copies =======================
        // rows >= M and columns >= K are zero-filled by the copy engine (tensor-map bounds), no tail code anywhere
        const bool leader = elect_one();
        long long tile_i = blockIdx.x;
        int kb = 0, rslot = 0, ruse = 0;
        bool ok = true;
#ifdef PU_TC_TIMELINE
        int tl_item = 0;
#endif
        while (tile_i < q.ntiles) {
            if (ruse >= 1) ok = mbar_wait(&raw_free[rslot], (uint32_t)((ruse - 1) & 1)) && ok;   // all converters have read the slot
#ifdef PU_TC_TIMELINE
            PU_TL(0, tl_item, 0);
            ++tl_item;
#endif
            if (leader) {
                mbar_expect_tx(&raw_full[rslot], (uint32_t)A_BYTES);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(raw_ring + (size_t)rslot * A_BYTES)), "l"(&tmap_a), "r"(kb * BK), "r"((int)(tile_i * BM)),
                               "r"(smem_u32(&raw_full[rslot]))
                             : "memory");
            }
            if (++kb == nkb) { kb = 0; tile_i += gridDim.x; }
            if (++rslot == D) { rslot = 0; ++ruse; }
        }
        if (!ok) s_err = 1;
    } else if (warp == (P_THREADS + E_THREADS) / 32 + 2) {
        // ======================= weight loader (streamed variant): one elected lane feeds the TA weight stages =======================
        // The per-role timeline (profiles/r1_tc_timeline_*.txt) showed the issuer spending ~1565 cycles per k-block of which
        // only ~690 issue MMAs: the rest was serial bookkeeping on the same warp, ~300 cycles of it the two single-lane
        // instructions of the weight fetch.  In a warp of its own the fetch runs up to TA items ahead and costs the issuer
        // nothing but the b_full wait.
        if constexpr (STREAM) {
            const bool leader = elect_one();
            long long f_tile = blockIdx.x;
            int f_kb = 0, f_slot = 0, f_use = 0;
            bool ok = true;
            while (f_tile < q.ntiles) {
                if (f_use >= 1) ok = mbar_wait(&stage_free[f_slot], (uint32_t)((f_use - 1) & 1)) && ok;  // previous reader retired
                if (leader) {
                    mbar_expect_tx(&b_full[f_slot], (uint32_t)B_KB);
                    bulk_g2s(b_res + (size_t)f_slot * B_KB, b_packed + (size_t)f_kb * B_KB, (uint32_t)B_KB, &b_full[f_slot]);
                }
                if (++f_kb == nkb) { f_kb = 0; f_tile += gridDim.x; }
                if (++f_slot == TA) { f_slot = 0; ++f_use; }
            }
            if (!ok) s_err = 1;
        }
    } else if (warp == (P_THREADS + E_THREADS) / 32) {
        // ======================= MMA issuer =======================
        // The whole warp walks the loop converged (all lanes poll the mbarriers); one elected lane issues.  Descriptors
        // are base + constant: start address field (bits 0-13, units of 16 B) never carries out for < 256 KB of smem.
        {
            const bool leader = elect_one();
            const uint32_t idesc = make_idesc(BN);
            const uint64_t bdesc0 = make_desc(smem_u32(b_res));
            long long cur_tile = blockIdx.x;
            int cur_kb = 0, slot = 0, use = 0, tile_count = 0;
            bool ok = true;
#ifdef PU_TC_TIMELINE
            int tl_item = 0;
#endif
            while (cur_tile < q.ntiles) {
                const bool last_kb = cur_kb == nkb - 1;
                const int buf = tile_count & 1, v = tile_count >> 1;
                if (cur_kb == 0 && v >= 1) ok = mbar_wait(&acc_empty[buf], (uint32_t)((v - 1) & 1)) && ok;
                ok = mbar_wait(&stage_ready[slot], (uint32_t)(use & 1)) && ok;           // all 256 producers filled the stage
#ifdef PU_TC_TIMELINE
                PU_TL(2, tl_item, 0);
#endif
                if constexpr (STREAM) ok = mbar_wait(&b_full[slot], (uint32_t)(use & 1)) && ok;  // this k-block of the weight landed
#ifdef PU_TC_TIMELINE
                PU_TL(2, tl_item, 1);
#endif
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                const uint32_t a_tmem = tmem_base + (uint32_t)(ACC_COLS + slot * 64);
                const uint64_t dbh0 = bdesc0 + (uint64_t)((STREAM ? slot : cur_kb) * (B_KB >> 4));
                const uint64_t dbl0 = dbh0 + (uint64_t)((BN * 128) >> 4);
                if (leader) {
#pragma unroll
                    for (int j = 0; j < BK / UMMA_K; ++j) {
                        umma_tf32_ts(d_tmem, a_tmem + j * UMMA_K, dbh0 + (uint64_t)(2 * j), idesc, (cur_kb > 0 || j > 0) ? 1u : 0u);
                        if (split) {
                            umma_tf32_ts(d_tmem, a_tmem + j * UMMA_K, dbl0 + (uint64_t)(2 * j), idesc, 1u);
                            umma_tf32_ts(d_tmem, a_tmem + 32 + j * UMMA_K, dbh0 + (uint64_t)(2 * j), idesc, 1u);
                        }
                    }
                    umma_commit(&stage_free[slot]);
                    if (last_kb) umma_commit(&acc_full[buf]);
                }
                int slot1 = slot + 1, use1 = use;
                if (slot1 == TA) { slot1 = 0; ++use1; }
#ifdef PU_TC_TIMELINE
                PU_TL(2, tl_item, 2);
#endif
#ifdef PU_TC_TIMELINE
                PU_TL(2, tl_item, 3);
                ++tl_item;
#endif
                if (++cur_kb == nkb) { cur_kb = 0; cur_tile += gridDim.x; tile_count++; }
                slot = slot1; use = use1;
            }
            if (!ok) s_err = 1;
        }
    } else {
        // ======================= epilogue: two independent groups of 8 warps =======================
        // Group g finishes the tiles whose accumulator lives in TMEM buffer g (every other tile of this CTA) with its own
        // staging tile and named barrier: while one group waits for its accumulator and drains tensor memory, the other
        // one does its arithmetic and stores -- the phases of consecutive tiles overlap instead of queueing behind
        // CTA-wide barriers.
        constexpr int G_THREADS = E_THREADS / 2;
        const int etid = tid - P_THREADS;
        const int eg = etid / G_THREADS, gtid = etid % G_THREADS, gwarp = gtid >> 5;
        const int quarter = gwarp & 3, chalf = gwarp >> 2;  // TMEM lane quarter (= warp id % 4) and column half
        float *tile_g = tile + (size_t)eg * BM * LDT;
        float *red_g = s_red + eg * (8 * 64 * 2);
        const int bar_id = 2 + eg;
        bool ok = true;
        constexpr int PAIRS = EPI == EPI_STORE ? 1 : ((BM / 16) * EC + G_THREADS - 1) / G_THREADS;  // (point, channel) pairs per thread
        constexpr int PP_ROWS = (G_THREADS / EC) * 16;  // tile rows between a thread's consecutive pairs
        // x_ring: pair (point pl, channel c) reads rows 16 pl + k of k-block c / 32; the swizzled byte offset of row 16 pl + k
        // is xo[k & 7] + 1024 (k >> 3)  (16 pl is a multiple of 8, so row & 7 == k & 7)
        uint32_t xo[8];
        {
            const int pl = gtid / EC, c = gtid % EC;
#pragma unroll
            for (int k = 0; k < 8; ++k) xo[k] = (uint32_t)((pl * 16 + k) * 128 + ((((c & 31) >> 2) ^ k) << 4) + (c & 3) * 4);
        }
        int eslot = eg * nkb, euse = 0;  // ring position of this group's current tile (x_ring only: nkb <= 2 < D)
        int tile_count = eg;
        for (long long tile_i = blockIdx.x + (long long)eg * gridDim.x; tile_i < q.ntiles; tile_i += 2LL * gridDim.x, tile_count += 2) {
            const int buf = eg, v = tile_count >> 1;
            const long long m0 = tile_i * BM;
            const long long rows_here = min((long long)BM, p.M - m0);
            int eslot1 = eslot + 1, euse1 = euse;  // second k-block of the tile
            if (eslot1 == D) { eslot1 = 0; ++euse1; }
#pragma unroll 1
            for (int pass = 0; pass < NPASS; ++pass) {
                const int nb = n0 + pass * EC;  // first global column of this pass
                // linear epilogue in accumulate mode: the old values of C do not depend on the MMA -- request them now, so
                // their HBM latency overlaps the accumulator wait and the TMEM read-out
                constexpr int CG = EC / 4, RLANES = G_THREADS / CG, RPT = BM / RLANES;  // 4-column groups, row lanes, rows per thread
                float4 old[EPI == EPI_STORE ? RPT : 1];
                if constexpr (EPI == EPI_STORE) {
                    const int gn = nb + (gtid % CG) * 4, rl = gtid / CG;
                    if (p.accumulate && gn + 3 < p.N && ((p.ldc & 3) == 0) && ((((uintptr_t)p.C) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < RPT; ++i) {
                            const int r = rl + i * RLANES;
                            old[i] = r < rows_here ? *reinterpret_cast<const float4 *>(p.C + (size_t)(m0 + r) * p.ldc + gn)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
#ifdef PU_TC_TIMELINE
                const int tl_item = (tile_count >> 1) * NPASS + pass;
                if (gwarp == 0) PU_TL(3 + eg, tl_item, 0);
#endif
                if (pass == 0) {
                    ok = mbar_wait(&acc_full[buf], (uint32_t)(v & 1)) && ok;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
#ifdef PU_TC_TIMELINE
                if (gwarp == 0) PU_TL(3 + eg, tl_item, 1);
#endif
                {   // TMEM -> registers -> staging tile: warp (quarter, chalf) moves 32 rows x EC/2 columns
                    const int row = quarter * 32 + lane;
#pragma unroll
                    for (int cc = 0; cc < EC / 2; cc += 16) {
                        const int c0 = chalf * (EC / 2) + cc;
                        float vals[16];
                        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + pass * EC + c0), vals);
#pragma unroll
                        for (int qd = 0; qd < 16; qd += 4)
                            *reinterpret_cast<float4 *>(&tile_g[row * LDT + c0 + qd]) =
                                make_float4(vals[qd], vals[qd + 1], vals[qd + 2], vals[qd + 3]);
                    }
                }
                if (pass == NPASS - 1 && !FUSED) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive_warp(&acc_empty[buf]);  // the tensor core may overwrite this accumulator now
                }   // (fused att backward: the buffer becomes the accumulator of the second MMA and is released after that)
                bar_sync_named(bar_id, G_THREADS);
#ifdef PU_TC_TIMELINE
                if (gwarp == 0) PU_TL(3 + eg, tl_item, 2);
#endif

                if constexpr (EPI == EPI_STORE) {
                    const bool vecC = ((p.ldc & 3) == 0) && ((((uintptr_t)p.C) & (p.c_bf16 ? 7 : 15)) == 0);
                    // thread -> fixed group of 4 columns, rows strided: the batch-norm partials accumulate in registers during
                    // the copy-out.  Shifted single pass: sums of (v - sh) and (v - sh)^2 with sh = the tile's first stored
                    // row, so M2 = S2 - S1^2/n loses nothing to cancellation.
                    const int cg = gtid % CG, rl = gtid / CG, c = cg * 4, gn = nb + c;
                    float sh[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
                    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.bias) {
                        bias4.x = gn + 0 < p.N ? p.bias[gn + 0] : 0.f; bias4.y = gn + 1 < p.N ? p.bias[gn + 1] : 0.f;
                        bias4.z = gn + 2 < p.N ? p.bias[gn + 2] : 0.f; bias4.w = gn + 3 < p.N ? p.bias[gn + 3] : 0.f;
                    }
                    if (p.stat_sum) {  // shift = stored value of row 0 (recomputed identically by every thread of the group)
                        const float4 t0 = *reinterpret_cast<float4 *>(&tile_g[c]);
                        sh[0] = t0.x + bias4.x; sh[1] = t0.y + bias4.y; sh[2] = t0.z + bias4.z; sh[3] = t0.w + bias4.w;
                    }
                    const bool fast = gn + 3 < p.N && vecC;
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        const int r = rl + i * RLANES;
                        if (r >= rows_here) continue;
                        float4 val = *reinterpret_cast<float4 *>(&tile_g[r * LDT + c]);
                        val.x += bias4.x; val.y += bias4.y; val.z += bias4.z; val.w += bias4.w;
                        float *cptr = p.C + (size_t)(m0 + r) * p.ldc + gn;
                        if (p.c_bf16) {   // bf16 storage of the pre-normalisation activations (statistics below: fp32 values)
                            unsigned short *cb = reinterpret_cast<unsigned short *>(p.C) + (size_t)(m0 + r) * p.ldc + gn;
                            if (fast) {
                                *reinterpret_cast<uint2 *>(cb) = make_uint2(pack_bf16x2(val.x, val.y), pack_bf16x2(val.z, val.w));
                            } else {
                                const float vv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (gn + j < p.N) cb[j] = (unsigned short)(pack_bf16x2(vv[j], 0.f) & 0xffffu);
                            }
                        } else if (fast) {
                            if (p.accumulate) { val.x += old[i].x; val.y += old[i].y; val.z += old[i].z; val.w += old[i].w; }
                            *reinterpret_cast<float4 *>(cptr) = val;
                        } else {
                            float vv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (gn + j < p.N) {
                                    if (p.accumulate) vv[j] += cptr[j];
                                    cptr[j] = vv[j];
                                }
                            val = make_float4(vv[0], vv[1], vv[2], vv[3]);
                        }
                        const float dv[4] = {val.x - sh[0], val.y - sh[1], val.z - sh[2], val.w - sh[3]};
#pragma unroll
                        for (int j = 0; j < 4; ++j) { s1[j] += dv[j]; s2[j] = fmaf(dv[j], dv[j], s2[j]); }
                    }
                    if (p.stat_sum) {
                        // a warp holds 32/CG consecutive row lanes of the same column groups: combine them with shuffles
                        // first (fixed order), so one partial per warp and column reaches shared memory
                        constexpr int RL2 = RLANES / (32 / CG);
                        static_assert(RL2 == G_THREADS / 32, "one partial per warp of the group");
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
#pragma unroll
                            for (int o = CG; o < 32; o <<= 1) {
                                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
                                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
                            }
                        }
                        if (lane < CG) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                red_g[(gwarp * EC + c + j) * 2 + 0] = s1[j];
                                red_g[(gwarp * EC + c + j) * 2 + 1] = s2[j];
                            }
                        }
                        bar_sync_named(bar_id, G_THREADS);
                        if (gtid < EC && nb + gtid < p.N) {
                            float a = 0.f, b = 0.f;
                            for (int l = 0; l < RL2; ++l) { a += red_g[(l * EC + gtid) * 2]; b += red_g[(l * EC + gtid) * 2 + 1]; }
                            const float shc = tile_g[gtid] + (p.bias ? p.bias[nb + gtid] : 0.f);  // same shift as above
                            const float n = (float)rows_here;
                            p.stat_sum[(size_t)tile_i * p.N + nb + gtid] = fmaf(n, shc, a);
                            p.stat_m2[(size_t)tile_i * p.N + nb + gtid] = fmaxf(b - a * a / n, 0.f);
                        }
                    }
                } else {
                    // one (point, channel) pair per thread step: the 16 neighbour rows of a point are consecutive rows of the tile
                    const long long pt0 = m0 / 16;
                    const int npts = (int)(rows_here / 16);
#pragma unroll
                    for (int pp = 0; pp < PAIRS; ++pp) {
                        const int pair = gtid + pp * G_THREADS;
                        const int pl = pair / EC, c = pair % EC;
                        const int gn = nb + c;
                        if (pair >= (BM / 16) * EC || pl >= npts || gn >= p.N) continue;
                        float a[16], x[16];
                        if (x_ring) {  // the tile's k-blocks are still resident in the TMA ring
                            const bool second = c >= 32;
                            ok = mbar_wait(&raw_full[second ? eslot1 : eslot], (uint32_t)((second ? euse1 : euse) & 1)) && ok;
                            const char *xs = raw_ring + (size_t)(second ? eslot1 : eslot) * A_BYTES + pp * (PP_ROWS * 128);
#pragma unroll
                            for (int k = 0; k < 16; ++k) x[k] = *reinterpret_cast<const float *>(xs + xo[k & 7] + (k >> 3) * 1024);
                        } else {
                            const float *xp = q.X + (size_t)(m0 + pl * 16) * q.ldx + gn;
#pragma unroll
                            for (int k = 0; k < 16; ++k) x[k] = xp[(size_t)k * q.ldx];
                        }
                        float g = 0.f;
                        if constexpr (ATT_BWD) g = q.G[(size_t)(pt0 + pl) * q.ldg + gn];
                        float mx = -FLT_MAX;
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            a[k] = tile_g[(pl * 16 + k) * LDT + c];
                            mx = fmaxf(mx, a[k]);
                        }
                        float sum = 0.f;
                        const float mxl = mx * 1.4426950408889634f;
#pragma unroll
                        for (int k = 0; k < 16; ++k) { a[k] = ex2_approx(fmaf(a[k], 1.4426950408889634f, -mxl)); sum += a[k]; }
                        const float inv = rcp_approx(sum);  // sum in [1, 16]
                        if constexpr (EPI == EPI_ATT_FWD) {
                            float num = 0.f;
#pragma unroll
                            for (int k = 0; k < 16; ++k) num = fmaf(x[k], a[k], num);
                            q.OUT[(size_t)(pt0 + pl) * q.ldo + gn] = num * inv;
                        } else {
                            float dot = 0.f;
#pragma unroll
                            for (int k = 0; k < 16; ++k) { a[k] *= inv; dot = fmaf(g * x[k], a[k], dot); }
                            float *cp = p.C + (size_t)(m0 + pl * 16) * p.ldc + gn;
                            float *op = q.OUT + (size_t)(m0 + pl * 16) * q.ldo + gn;
                            if constexpr (FUSED) {   // d_act also goes back into the staging tile: operand of the second MMA
#pragma unroll
                                for (int k = 0; k < 16; ++k) tile_g[(pl * 16 + k) * LDT + c] = a[k] * (g * x[k] - dot);
                            }
                            if (p.ldc == BN && q.ldo == BN) {  // contiguous outputs: immediate offsets
#pragma unroll
                                for (int k = 0; k < 16; ++k) {
                                    cp[k * BN] = a[k] * (g * x[k] - dot);   // d_act
                                    op[k * BN] = g * a[k];                  // dx_direct
                                }
                            } else {
#pragma unroll
                                for (int k = 0; k < 16; ++k) {
                                    cp[(size_t)k * p.ldc] = a[k] * (g * x[k] - dot);
                                    op[(size_t)k * q.ldo] = g * a[k];
                                }
                            }
                        }
                    }
                }
                if constexpr (FUSED) {
                    // ---- dx += d_act w^T on the tensor core, inside this epilogue ----
                    // tile_count is the CTA-local tile number; the single A2 operand buffer is used in that order by the two
                    // groups: tile t may write it once the second MMA of tile t - 1 has retired (a2_free phase t - 1).
                    bar_sync_named(bar_id, G_THREADS);                     // d_act tile complete, dx_direct stores issued
                    if (tile_count >= 1) ok = mbar_wait(&a2_free, (uint32_t)((tile_count - 1) & 1)) && ok;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    {   // rows of d_act -> hi/lo -> tensor memory: warp (quarter, chalf) = 32 rows x 32 columns
                        const int row = quarter * 32 + lane;
                        const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(A2_COL + chalf * 32);
#pragma unroll
                        for (int cc = 0; cc < 32; cc += 16) {
                            float vv[16], hh[16];
#pragma unroll
                            for (int qd = 0; qd < 16; qd += 4) {
                                const float4 t4 = *reinterpret_cast<const float4 *>(&tile_g[row * LDT + chalf * 32 + cc + qd]);
                                vv[qd] = t4.x; vv[qd + 1] = t4.y; vv[qd + 2] = t4.z; vv[qd + 3] = t4.w;
                            }
                            if (split) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) { hh[i] = tf32_rn(vv[i]); vv[i] -= hh[i]; }
                                tmem_st16(ta + cc, hh);
                                tmem_st16(ta + BN + cc, vv);
                            } else {
                                tmem_st16(ta + cc, vv);
                            }
                        }
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    bar_sync_named(bar_id, G_THREADS);                     // operand complete
                    if (gwarp == 0) {                                      // converged warp, one elected lane issues
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const bool leader2 = elect_one();
                        const uint32_t idesc2 = make_idesc(BN);
                        const uint64_t b2desc0 = make_desc(smem_u32(b2_res));
                        const uint32_t d2 = tmem_base + (uint32_t)(buf * BN);      // the drained accumulator of this tile
                        const uint32_t a2 = tmem_base + (uint32_t)A2_COL;
                        if (leader2) {
#pragma unroll
                            for (int kk = 0; kk < BN / UMMA_K; ++kk) {             // contraction over the BN channels of d_act
                                const uint64_t bh = b2desc0 + (uint64_t)((kk >> 2) * (B_KB >> 4)) + (uint64_t)(2 * (kk & 3));
                                const uint64_t bl = bh + (uint64_t)((BN * 128) >> 4);
                                umma_tf32_ts(d2, a2 + kk * UMMA_K, bh, idesc2, kk > 0 ? 1u : 0u);
                                if (split) {
                                    umma_tf32_ts(d2, a2 + kk * UMMA_K, bl, idesc2, 1u);
                                    umma_tf32_ts(d2, a2 + BN + kk * UMMA_K, bh, idesc2, 1u);
                                }
                            }
                            umma_commit(&a2_free);
                            umma_commit(&acc2_full[buf]);
                        }
                    }
                    {   // dx = dx_direct (stored above, before the group barrier) + d_act w^T: lane = row, 32 columns per warp.
                        // The sum is formed by 16-byte fire-and-forget reductions in L2 (one per address: deterministic);
                        // loading the old values back instead costs 0.19 ms per launch at 2.88 M rows.
                        const int row = quarter * 32 + lane;
                        float *op = q.OUT + (size_t)(m0 + row) * q.ldo + n0 + chalf * 32;
                        ok = mbar_wait(&acc2_full[buf], (uint32_t)(v & 1)) && ok;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                        for (int cc = 0; cc < 32; cc += 16) {
                            float vals[16];
                            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + chalf * 32 + cc), vals);
                            if (row < rows_here) {
#pragma unroll
                                for (int qd = 0; qd < 16; qd += 4)
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + cc + qd), "f"(vals[qd]),
                                                 "f"(vals[qd + 1]), "f"(vals[qd + 2]), "f"(vals[qd + 3]) : "memory");
                            }
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive_warp(&acc_empty[buf]);  // now the first MMA of tile + 2 may overwrite the buffer
                }
                bar_sync_named(bar_id, G_THREADS);  // the staging tile is reused by the next pass / tile of this group
#ifdef PU_TC_TIMELINE
                if (gwarp == 0) PU_TL(3 + eg, tl_item, 3);
#endif
            }
            if (x_ring) {  // every x value has been consumed by the arithmetic above: the loader may refill the tile's slots
                mbar_arrive_warp(&raw_free[eslot]);
                if (nkb == 2) mbar_arrive_warp(&raw_free[eslot1]);
                eslot += 2 * nkb;  // the other group owns the tile in between
                if (eslot >= D) { eslot -= D; ++euse; }
            }
        }
        if (!ok) s_err = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
    if (tid == 0 && s_err && p.error_flag) *p.error_flag = 1;
}

constexpr size_t kMaxDynSmem = 227 * 1024 - 10 * 1024;  // dynamic budget: 227 KB minus the kernel's static shared memory (~9 KB)
template <int BN, bool STREAM>
static size_t persist_fixed_bytes(int K, bool second_weight = false) {  // everything except the raw ring
    const int nkb = (K + BK - 1) / BK;
    const int ec = BN > 64 ? 64 : BN;
    const int ta = BN <= 64 ? MAX_TA : 3;
    return (size_t)(STREAM ? ta : nkb) * 2 * BN * 128 * (second_weight ? 2 : 1) + 2 * (size_t)BM * (ec + 4) * 4 + 1024;  // weights, 2 staging tiles
}
template <int BN, bool STREAM>
static int persist_raw_depth(int K, bool second_weight = false) {  // 0 => does not fit
    const size_t fixed = persist_fixed_bytes<BN, STREAM>(K, second_weight);
    if (fixed + 3 * (size_t)BM * 128 > kMaxDynSmem) return 0;
    long long d = (long long)((kMaxDynSmem - fixed) / ((size_t)BM * 128));
    return (int)(d > MAX_RAW_P ? MAX_RAW_P : d);
}

static inline size_t packed_weight_bytes(int K, int N, int bn) {
    const int nkb = (K + BK - 1) / BK, nslabs = (N + bn - 1) / bn;
    return (size_t)nslabs * nkb * 2 * bn * 128;
}

// 2-D tensor map of the row-major operand x[M, K] (row stride lda floats): box = 128 rows x 32 floats, 128-byte swizzle,
// out-of-bounds elements read as zero.  cuTensorMapEncodeTiled is a pure host-side encoder; it is looked up through the
// runtime (cudaGetDriverEntryPoint), so the library has no link-time dependency on libcuda.
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (TensorMapEncodeFn)p;
    }();
    return fn;
}
static int make_tmap_rows(CUtensorMap *map, const float *base, long long M, int K, int lda, int box_rows) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return PU_ERR_UNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)lda * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PU_OK : PU_ERR_INVALID_ARG;
}

#ifdef PU_TC_TIMELINE
static long long *g_timeline = nullptr;
extern "C" __attribute__((visibility("default"))) void pu_tc_debug_set_timeline(long long *device_buffer) { g_timeline = device_buffer; }
extern "C" __attribute__((visibility("default"))) int pu_tc_debug_timeline_dims(int *roles, int *items, int *events) {
    *roles = TL_ROLES; *items = TL_ITEMS; *events = TL_EVENTS;
    return 0;
}
#endif

template <int BN, int EPI, bool STREAM>
static int launch_persist(const Params2 &q, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    Params2 qq = q;
#ifdef PU_TC_TIMELINE
    qq.timeline = g_timeline;
#endif
    constexpr bool kSecondWeight = EPI == EPI_ATT_BWD_F;
    qq.raw_depth = persist_raw_depth<BN, STREAM>(q.g.K, kSecondWeight);
    if (qq.raw_depth < 3) return PU_ERR_UNSUPPORTED;
    const size_t smem = persist_fixed_bytes<BN, STREAM>(q.g.K, kSecondWeight) + (size_t)qq.raw_depth * BM * 128;
    static bool configured = false;
    if (!configured) {
        PU_CUDA_TRY(cudaFuncSetAttribute(tc_persist_kernel<BN, EPI, STREAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
        configured = true;
    }
    const int ny = ceil_div(q.g.N, BN);
    if (STREAM) {
        const size_t need = packed_weight_bytes(q.g.K, q.g.N, BN);
        if (!workspace || workspace_bytes < need || (((uintptr_t)workspace) & 15)) return PU_ERR_WORKSPACE;
        const int nkb = (q.g.K + BK - 1) / BK;
        const long long chunks = (long long)ny * nkb * BN * 8;
        tc_pack_weight_kernel<<<(unsigned)((chunks + 255) / 256 > 4096 ? 4096 : (chunks + 255) / 256), 256, 0, st>>>(
            q.g.Bt, q.g.ldb, q.g.N, q.g.K, BN, nkb, ny, q.g.mode == 3 ? 1 : 0, (char *)workspace);
        PU_LAUNCH_CHECK();
        qq.Bp = (const char *)workspace;
    }
    long long gx = kNumSMs / ny;  // ~one CTA per SM in total (227 KB-class shared memory); each walks its row tiles
    if (gx < 1) gx = 1;
    if (gx > q.ntiles) gx = q.ntiles;
    dim3 grid((unsigned)gx, ny);
    CUtensorMap tmap;
    const int trc = make_tmap_rows(&tmap, q.g.A, q.g.M, q.g.K, q.g.lda, BM);
    if (trc != PU_OK) return trc;
    tc_persist_kernel<BN, EPI, STREAM><<<grid, PERSIST_THREADS, smem, st>>>(qq, tmap);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

// Shape policy.  Resident weights (no packing pass) when the whole [K x BN] hi/lo pair fits beside a >= 3-deep raw ring
// with BN covering N in at most two slabs; otherwise the streamed-weight variant with the widest tile (BN = 128).
struct Choice { int bn; bool stream; };
static Choice choose_linear(int K, int N) {
    if (N <= 32 && persist_raw_depth<32, false>(K) >= 5) return {32, false};
    if (N <= 64 && persist_raw_depth<64, false>(K) >= 5) return {64, false};
    if (N > 64 && persist_raw_depth<128, false>(K) >= 5) return {128, false};
    if (N > 64 && N <= 128 && persist_raw_depth<64, false>(K) >= 5) return {64, false};
    if (N <= 32) return {32, true};
    if (N <= 64) return {64, true};
    return {128, true};
}
static Choice choose_att(int d) {
    if (d <= 32 && persist_raw_depth<32, false>(d) >= 5) return {32, false};
    if (d <= 64 && persist_raw_depth<64, false>(d) >= 5) return {64, false};
    return {128, true};
}
static size_t tc_workspace_bytes(int K, int N) {
    size_t m = packed_weight_bytes(K, N, 128);
    const size_t a = packed_weight_bytes(K, N, 64), b = packed_weight_bytes(K, N, 32);
    if (a > m) m = a;
    if (b > m) m = b;
    return m + 256;
}

template <int EPI>
static int dispatch_persist(const Params2 &q, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    const Choice ch = EPI == EPI_STORE ? choose_linear(q.g.K, q.g.N) : choose_att(q.g.K);
    if (!ch.stream) {
        switch (ch.bn) {
            case 32: return launch_persist<32, EPI, false>(q, workspace, workspace_bytes, st);
            case 64: return launch_persist<64, EPI, false>(q, workspace, workspace_bytes, st);
            case 128: return launch_persist<128, EPI, false>(q, workspace, workspace_bytes, st);
        }
    } else {
        switch (ch.bn) {
            case 32: return launch_persist<32, EPI, true>(q, workspace, workspace_bytes, st);
            case 64: return launch_persist<64, EPI, true>(q, workspace, workspace_bytes, st);
            case 128: return launch_persist<128, EPI, true>(q, workspace, workspace_bytes, st);
        }
    }
    return PU_ERR_UNSUPPORTED;
}


// =============================================================================================================
// Tensor-core weight gradient:  dW[Kin, N] = sum_r X[r, Kin]^T dY[r, N]   (+ db = column sums of dY via a ones row)
// The reduction runs over ROWS, so both operands are "MN-major" for the MMA (element (m, k) of A is X[k][m]: m
// contiguous).  For 32-bit operands the only MN-major shared-memory layout the tensor core accepts is
// SWIZZLE_128B_BASE32B (cute::UMMA::Layout_MN_SW128_32B_Atom): a 32-column slab of a row-major matrix is stored as rows
// of 128 B, and inside every 512-byte group of 4 rows the 32-BYTE chunk index is XOR-ed with (row % 4)
// (Swizzle<2,5,2>).  Descriptors: layout type 1, LBO = slab stride, SBO = stride between 4-row groups (512 B).  Each CTA owns a contiguous chunk of rows and one (128 x BN) block of dW, accumulates ALL its rows
// in TMEM and writes one partial at the end; pu_tc_wgrad then reduces the partials in fixed order (deterministic).
struct WParams {
    const float *X; int ldx;   // [M, Kin]
    const float *G; int ldg;   // [M, N]
    long long M; int Kin, N;
    long long rows_per_cta;
    float *part;               // [gridDim.x][Kin][N]
    float *db_part;            // [gridDim.x][N] or null (needs Kin % 128 != 0 for the free ones row)
    int mode;
    int raw_depth;
    int *error_flag;
};

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t slab_bytes) {
    // MN-major SWIZZLE_128B_BASE32B: LBO = byte stride between 32-element MN slabs, SBO = byte stride between 4-row k groups
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((slab_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// byte offset inside a slab of the 16-byte piece holding columns cin..cin+3 (cin % 4 == 0, < 32) of k-row r
__device__ __forceinline__ uint32_t sw128_32b(int r, int cin) {
    return (uint32_t)((r >> 2) * 512 + (r & 3) * 128 + ((((cin >> 3) ^ (r & 3)) & 3) << 5) + ((cin >> 2) & 1) * 16);
}
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {  // as make_idesc, a_major = b_major = 1 (MN-major)
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

constexpr int WG_THREADS = 512;          // 16 producer warps
constexpr int WG_TOTAL = WG_THREADS + 32;  // + 1 MMA-issuer warp
constexpr int WG_ROWS = 32;  // rows (= MMA k) per pipeline step: 4 instructions of K = 8

// copies a [32 rows x COLS] block of a row-major matrix (columns col0.., zero beyond ncols / row_end) into a raw slot
template <int COLS>
__device__ __forceinline__ void wg_raw_issue(char *slot, const float *__restrict__ base, int ld, long long row0, long long row_end,
                                             int col0, int ncols, int tid) {
    constexpr int CHUNKS = WG_ROWS * COLS / 4;  // 16-byte chunks
#pragma unroll
    for (int i = 0; i < (CHUNKS + WG_THREADS - 1) / WG_THREADS; ++i) {
        const int idx = tid + WG_THREADS * i;
        if (idx < CHUNKS) {
            const int r = idx / (COLS / 4), c4 = (idx % (COLS / 4)) * 4;
            const long long gr = row0 + r;
            const bool valid = gr < row_end && col0 + c4 < ncols;  // ncols % 4 == 0 (host)
            const float *src = valid ? base + (size_t)gr * ld + col0 + c4 : base;
            cp_async16(slot + (size_t)idx * 16, src, valid);
        }
    }
}
// raw [32][COLS] -> operand image: slab s = 32 columns, each slab 32 rows x 128 B swizzled; optional ones row at column
// `ones_col` (the bias-gradient trick) for valid rows
template <int COLS>
__device__ __forceinline__ void wg_convert(const char *slot, char *hi, char *lo, int tid, bool split, int ones_col,
                                           long long row0, long long row_end) {
    constexpr int CHUNKS = WG_ROWS * COLS / 4;
#pragma unroll
    for (int i = 0; i < (CHUNKS + WG_THREADS - 1) / WG_THREADS; ++i) {
        const int idx = tid + WG_THREADS * i;
        if (idx < CHUNKS) {
            const int r = idx / (COLS / 4), c4 = (idx % (COLS / 4)) * 4;
            float4 v = *reinterpret_cast<const float4 *>(slot + (size_t)idx * 16);
            if (ones_col >= c4 && ones_col < c4 + 4 && row0 + r < row_end) {
                const float one = 1.f;
                if (ones_col == c4) v.x = one; else if (ones_col == c4 + 1) v.y = one; else if (ones_col == c4 + 2) v.z = one; else v.w = one;
            }
            const uint32_t off = (uint32_t)((c4 >> 5) * (WG_ROWS * 128)) + sw128_32b(r, c4 & 31);
            if (split) {
                const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
                *reinterpret_cast<float4 *>(hi + off) = h;
                *reinterpret_cast<float4 *>(lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            } else {
                *reinterpret_cast<float4 *>(hi + off) = v;
            }
        }
    }
}

template <int BN>
__global__ void __launch_bounds__(WG_TOTAL, 1) tc_wgrad_kernel(const WParams w) {
    constexpr int A_BYTES = WG_ROWS * BM * 4;   // 16 KB: 4 slabs of 32 rows x 128 B
    constexpr int B_BYTES = WG_ROWS * BN * 4;
    constexpr int STAGE = 2 * (A_BYTES + B_BYTES);  // hi + lo of both operands
    constexpr int RAW = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ __align__(1024) char smem_raw[];
    char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // aligned by offset: stays a shared pointer
    __shared__ uint64_t stage_free[2], stage_ready[2], all_done;
    __shared__ uint32_t tmem_base_slot;
    __shared__ int s_err;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.y * BM;   // first Kin column (= MMA M index) of this CTA's block
    const int n0 = blockIdx.z * BN;
    const bool split = w.mode == 3;
    const int D = w.raw_depth;
    char *op_ring = smem;                       // 2 operand stages
    char *raw_ring = smem + 2 * STAGE;          // D raw slots
    const long long r_begin = (long long)blockIdx.x * w.rows_per_cta;
    const long long r_end = min(w.M, r_begin + w.rows_per_cta);
    const int nsteps = r_end > r_begin ? (int)((r_end - r_begin + WG_ROWS - 1) / WG_ROWS) : 0;
    // ones row: first padded M index of the last Kin block, if any
    const int ones_col = (w.db_part && blockIdx.y == gridDim.y - 1 && (w.Kin % BM) != 0) ? (w.Kin - k0) : -1;

    if (tid == 0) {
        mbar_init(&stage_free[0], 1); mbar_init(&stage_free[1], 1); mbar_init(&all_done, 1);
        mbar_init(&stage_ready[0], WG_THREADS / ARRIVE_DIV); mbar_init(&stage_ready[1], WG_THREADS / ARRIVE_DIV);
        s_err = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the operand stages start as zeros: the padding columns (Kin - k0 < 128, N - n0 < BN) are never written again
    for (int i = tid; i < 2 * STAGE / 16; i += WG_TOTAL) reinterpret_cast<float4 *>(op_ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t idesc = make_idesc_mn(BN);

    bool ok = true;
    if (warp < WG_THREADS / 32 && ones_col >= 0) {
        // ======================= producers, generic path (the CTA that also carries the ones row for db) =======================
        int issued = 0;
        auto issue_one = [&]() {
            if (issued < nsteps) {
                char *slot = raw_ring + (size_t)(issued % D) * RAW;
                const long long row0 = r_begin + (long long)issued * WG_ROWS;
                wg_raw_issue<BM>(slot, w.X, w.ldx, row0, r_end, k0, w.Kin, tid);
                wg_raw_issue<BN>(slot + A_BYTES, w.G, w.ldg, row0, r_end, n0, w.N, tid);
            }
            cp_async_commit();
            ++issued;
        };
        for (int i = 0; i < D - 1; ++i) issue_one();
        for (int it = 0; it < nsteps; ++it) {
            issue_one();
            cp_async_wait_dyn(D - 1);
            const int s = it & 1, u = it >> 1;
            char *a_hi = op_ring + (size_t)s * STAGE, *a_lo = a_hi + A_BYTES, *b_hi = a_lo + A_BYTES, *b_lo = b_hi + B_BYTES;
            if (u >= 1) ok = mbar_wait(&stage_free[s], (uint32_t)((u - 1) & 1)) && ok;
            const char *slot = raw_ring + (size_t)(it % D) * RAW;
            const long long row0 = r_begin + (long long)it * WG_ROWS;
            wg_convert<BM>(slot, a_hi, a_lo, tid, split, ones_col, row0, r_end);
            wg_convert<BN>(slot + A_BYTES, b_hi, b_lo, tid, split, -1, row0, r_end);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_warp(&stage_ready[s]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (warp < WG_THREADS / 32) {
        // ======================= producers =======================
        // Only the REAL columns are moved: chunk idx < nA belongs to X (row idx / cprA), the rest to dY.  Every thread
        // keeps its (<= MAXC) chunks in a small register table built once -- source pointer at the CTA's first row,
        // operand-image offset -- so a pipeline step costs one pointer add, one cp.async, one LDS and two STS per chunk.
        constexpr int MAXC = (WG_ROWS * (BM + BN) / 4 + WG_THREADS - 1) / WG_THREADS;
        const int colsA = min(w.Kin - k0, BM), colsB = min(w.N - n0, BN);
        const int cprA = colsA >> 2, cprB = colsB >> 2, nA = WG_ROWS * cprA, nT = nA + WG_ROWS * cprB;
        const float *c_src[MAXC];
        size_t c_step[MAXC];
        uint32_t c_op[MAXC], c_lo[MAXC];
        int c_row[MAXC];
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int idx = tid + WG_THREADS * i;
            c_src[i] = w.X; c_step[i] = 0; c_op[i] = 0; c_lo[i] = 0; c_row[i] = 0;
            if (idx < nA) {
                const int r = idx / cprA, c4 = (idx - r * cprA) * 4;
                c_src[i] = w.X + (size_t)(r_begin + r) * w.ldx + k0 + c4;
                c_step[i] = (size_t)WG_ROWS * w.ldx;
                c_op[i] = (uint32_t)((c4 >> 5) * (WG_ROWS * 128)) + sw128_32b(r, c4 & 31);
                c_lo[i] = A_BYTES; c_row[i] = r;
            } else if (idx < nT) {
                const int j = idx - nA, r = j / cprB, c4 = (j - r * cprB) * 4;
                c_src[i] = w.G + (size_t)(r_begin + r) * w.ldg + n0 + c4;
                c_step[i] = (size_t)WG_ROWS * w.ldg;
                c_op[i] = (uint32_t)(2 * A_BYTES + (c4 >> 5) * (WG_ROWS * 128)) + sw128_32b(r, c4 & 31);
                c_lo[i] = B_BYTES; c_row[i] = r;
            }
        }
        const uint32_t thr_raw = (uint32_t)tid * 16u;
        int issued = 0, iss_slot = 0, cur_slot = 0;
        auto issue_one = [&]() {
            if (issued < nsteps) {
                const uint32_t slot = smem_u32(raw_ring + (size_t)iss_slot * RAW) + thr_raw;
                const int rows_left = (int)min((long long)WG_ROWS, r_end - (r_begin + (long long)issued * WG_ROWS));
#pragma unroll
                for (int i = 0; i < MAXC; ++i) {
                    if (tid + WG_THREADS * i < nT) {
                        const int sz = c_row[i] < rows_left ? 16 : 0;  // rows past the CTA's chunk are zero-filled
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slot + i * (WG_THREADS * 16)),
                                     "l"(sz ? c_src[i] : w.X), "r"(sz) : "memory");
                        c_src[i] += c_step[i];
                    }
                }
            }
            cp_async_commit();
            ++issued;
            if (++iss_slot == D) iss_slot = 0;
        };
        for (int i = 0; i < D - 1; ++i) issue_one();
        for (int it = 0; it < nsteps; ++it) {
            issue_one();
            cp_async_wait_dyn(D - 1);
            const int s = it & 1, u = it >> 1;
            char *st_base = op_ring + (size_t)s * STAGE;
            if (u >= 1) ok = mbar_wait(&stage_free[s], (uint32_t)((u - 1) & 1)) && ok;
            const char *slot = raw_ring + (size_t)cur_slot * RAW + thr_raw;
#pragma unroll
            for (int i = 0; i < MAXC; ++i) {
                if (tid + WG_THREADS * i < nT) {
                    const float4 v = *reinterpret_cast<const float4 *>(slot + i * (WG_THREADS * 16));
                    char *dst = st_base + c_op[i];
                    if (split) {
                        const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
                        *reinterpret_cast<float4 *>(dst) = h;
                        *reinterpret_cast<float4 *>(dst + c_lo[i]) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                    } else {
                        *reinterpret_cast<float4 *>(dst) = v;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_warp(&stage_ready[s]);
            if (++cur_slot == D) cur_slot = 0;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        // ======================= MMA issuer: converged warp, one elected lane issues (see elect_one) =======================
        const bool leader = elect_one();
        const uint64_t adesc0 = make_desc_mn(smem_u32(op_ring), WG_ROWS * 128);
        for (int it = 0; it < nsteps; ++it) {
            const int s = it & 1, u = it >> 1;
            ok = mbar_wait(&stage_ready[s], (uint32_t)(u & 1)) && ok;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t dah0 = adesc0 + (uint64_t)((s * STAGE) >> 4), dal0 = dah0 + (uint64_t)(A_BYTES >> 4);
            const uint64_t dbh0 = dal0 + (uint64_t)(A_BYTES >> 4), dbl0 = dbh0 + (uint64_t)(B_BYTES >> 4);
            if (leader) {
#pragma unroll
                for (int j = 0; j < WG_ROWS / UMMA_K; ++j) {  // 8 rows = two 512-byte k groups per slab
                    const uint64_t o = (uint64_t)((j * 1024) >> 4);
                    umma_tf32(tmem_base, dah0 + o, dbh0 + o, idesc, (it > 0 || j > 0) ? 1u : 0u);
                    if (split) {
                        umma_tf32(tmem_base, dah0 + o, dbl0 + o, idesc, 1u);
                        umma_tf32(tmem_base, dal0 + o, dbh0 + o, idesc, 1u);
                    }
                }
                umma_commit(&stage_free[s]);
                if (it == nsteps - 1) umma_commit(&all_done);
            }
        }
    }
    if (warp < 8 && nsteps > 0) ok = mbar_wait(&all_done, 0u) && ok;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) s_err = 1;
    // epilogue: TMEM lane = Kin index (k0 + lane), column = N index; warps 0-3 and 4-7 split the columns
    if (warp < 8) {
        const int quarter = warp & 3, chalf = warp >> 2;
        const int m = quarter * 32 + lane;      // row of the dW block
        const int gk = k0 + m;
        float *prow = w.part + ((size_t)blockIdx.x * w.Kin + gk) * w.N + n0;
#pragma unroll
        for (int cc = 0; cc < BN / 2; cc += 16) {
            const int c0 = chalf * (BN / 2) + cc;
            float vals[16];
            if (nsteps > 0) tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, vals);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) vals[i] = 0.f;
            }
            if (gk < w.Kin) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (n0 + c0 + i < w.N) prow[c0 + i] = vals[i];
            } else if (m == ones_col && w.db_part) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (n0 + c0 + i < w.N) w.db_part[(size_t)blockIdx.x * w.N + n0 + c0 + i] = vals[i];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
    if (tid == 0 && s_err && w.error_flag) *w.error_flag = 1;
}

struct WgPlan { int bn, gx, gy, gz, depth; long long rows_per_cta; size_t smem; };
static WgPlan wgrad_plan_tc(long long M, int Kin, int N) {
    WgPlan pl{};
    pl.bn = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    pl.gy = ceil_div(Kin, BM);
    pl.gz = ceil_div(N, pl.bn);
    long long gx = kNumSMs / (pl.gy * pl.gz);
    if (gx < 1) gx = 1;
    const long long max_gx = (M + 4 * WG_ROWS - 1) / (4 * WG_ROWS);  // at least 128 rows per CTA
    if (gx > max_gx) gx = max_gx;
    if (gx < 1) gx = 1;
    long long rpc = (M + gx - 1) / gx;
    rpc = (rpc + 63) / 64 * 64;   // whole steps of either kernel (32- or 64-row steps): a CTA never reads its neighbour's rows
    pl.rows_per_cta = rpc;
    pl.gx = (int)((M + rpc - 1) / rpc);
    const size_t stage = 2 * ((size_t)WG_ROWS * BM * 4 + (size_t)WG_ROWS * pl.bn * 4);
    const size_t raw = (size_t)WG_ROWS * BM * 4 + (size_t)WG_ROWS * pl.bn * 4;
    long long d = (long long)((kMaxDynSmem - 2 * stage - 1024) / raw);
    pl.depth = (int)(d > MAX_RAW ? MAX_RAW : d);
    pl.smem = 2 * stage + (size_t)pl.depth * raw + 1024;
    return pl;
}

template <int BN>
static int launch_wgrad(const WParams &w, const WgPlan &pl, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        PU_CUDA_TRY(cudaFuncSetAttribute(tc_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
        configured = true;
    }
    dim3 grid(pl.gx, pl.gy, pl.gz);
    tc_wgrad_kernel<BN><<<grid, WG_TOTAL, pl.smem, st>>>(w);
    PU_LAUNCH_CHECK();
    return PU_OK;
}


// =============================================================================================================
// Weight gradient, second generation: the x operand goes through TENSOR MEMORY (TS-form MMA), only dy stays in shared memory.
// The first kernel (tc_wgrad_kernel above) keeps hi/lo images of BOTH operands in shared memory and is bound by that
// memory's 128 B/clk port: per 32-row step raw in + LDS + 2 STS + three operand reads by the tensor core come to ~7 bytes of
// shared-memory traffic per input byte (224 KB per 32 KB of input at BN = 128, 1750 cycles) -- 2.2-3.5 TB/s measured, and
// worse when Kin < 128 pads the A tile.  Here
//   * a loader warp brings [32 rows x boxA] of x and [32 x BN] of dy per step with two TMA tile copies (no swizzle; columns
//     beyond Kin / N and rows beyond M arrive as zeros, so there is no tail code),
//   * 8 "A" warps read x COLUMN-wise (lane = channel, 16 rows each: conflict-free 128-byte row segments), split hi/lo in
//     registers and tcgen05.st the values into an operand-A stage in tensor memory (lane = channel = MMA row m, column = k),
//   * 8 "B" warps convert dy into the MN-major SWIZZLE_128B_BASE32B hi/lo images as before,
//   * the issuer runs 12 TS-form MMAs per step (A from TMEM, B from shared memory).
// Shared-memory traffic drops to ~4.5 bytes per input byte at BN = 128 and does not grow when Kin is padded.
// Work split (second lesson of the per-role timeline, profiles/r2_wgrad_timeline_*.txt): these kernels are INSTRUCTION-ISSUE
// bound, and most instructions are per-step bookkeeping (two barrier waits, fences, arrivals, ring arithmetic: ~150 per warp
// and step) rather than conversion work (~700 warp instructions per 16 KB).  So a step is handled by as FEW warps as the
// tensor-memory lane rule allows -- four "A" warps (one per lane quarter, a thread converts all rows of its channel) and four
// "B" warps -- and there are TWO such groups per operand that take alternate steps, so two steps are always in conversion.
// Narrow operand pairs (x block and dy block <= 64 columns) use 64-row steps, which halves the bookkeeping per byte again.
constexpr int WT_GROUP = 128;                                  // threads per converter group (4 warps)
constexpr int WT_A_THREADS = 2 * WT_GROUP, WT_B_THREADS = 2 * WT_GROUP;
constexpr int WT_TOTAL = WT_A_THREADS + WT_B_THREADS + 64;   // + issuer warp + TMA warp
constexpr int WT_MAX_TA = 4;                                 // operand-A stages in tensor memory
constexpr int WT_MAX_RAW = 8;

struct WTParams {
    long long M; int Kin, N;
    long long rows_per_cta;
    float *part;               // [gridDim.x][Kin][N]
    float *db_part;            // [gridDim.x][N] or null
    int mode;
    int raw_depth;
    int box_a;                 // columns of x per step (32 / 64 / 128)
    int *error_flag;
    long long *timeline;       // optional (tools/wgrad_timeline.py): clock64 stamps of CTA 0, [role][step][event]
};
constexpr int WTL_STEPS = 96, WTL_EVENTS = 4, WTL_ROLES = 4;   // roles: 0 A converter (warp 0), 1 B converter, 2 issuer, 3 loader
#define WTL(role, step, ev)                                                                                         \
    do {                                                                                                            \
        if (w.timeline && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0 && (step) < WTL_STEPS) \
            w.timeline[((size_t)(role) * WTL_STEPS + (step)) * WTL_EVENTS + (ev)] = clock64();                      \
    } while (0)

__device__ __forceinline__ uint32_t make_idesc_ts_mn(int n) {  // A from tensor memory (K-major), B MN-major in shared memory
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int BN, int ROWS>
__global__ void __launch_bounds__(WT_TOTAL, 1) tc_wgrad_ts_kernel(const WTParams w, const __grid_constant__ CUtensorMap tmap_x,
                                                                  const __grid_constant__ CUtensorMap tmap_g) {
    constexpr int B_BYTES = ROWS * BN * 4;             // one image of the dy operand
    constexpr int B_STAGE = 2 * B_BYTES;               // hi + lo
    constexpr int ACC_COLS = BN < 32 ? 32 : BN;
    constexpr int A_COLS = 2 * ROWS;                   // tensor-memory columns of one operand-A stage: hi rows, then lo rows
    constexpr int TA = (512 - ACC_COLS) / A_COLS < WT_MAX_TA ? (512 - ACC_COLS) / A_COLS : WT_MAX_TA;
    constexpr int TMEM_COLS = 512;
    static_assert(TA >= 2 && ACC_COLS + TA * A_COLS <= TMEM_COLS, "tensor memory budget");
    extern __shared__ __align__(1024) char smem_raw[];
    char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t raw_full[WT_MAX_RAW], raw_free[WT_MAX_RAW], a_ready[WT_MAX_TA], a_free[WT_MAX_TA], b_ready[2], b_free[2], all_done;
    __shared__ uint32_t tmem_base_slot;
    __shared__ int s_err;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.y * BM, n0 = blockIdx.z * BN;
    const bool split = w.mode == 3;
    const int D = w.raw_depth, box_a = w.box_a;
    const int a_raw_bytes = ROWS * box_a * 4, raw_bytes = a_raw_bytes + B_BYTES;
    char *b_ring = smem;                        // 2 operand stages of dy (stage g belongs to B group g)
    char *raw_ring = smem + 2 * B_STAGE;        // D raw slots: [ROWS][box_a] of x, then [ROWS][BN] of dy
    const long long r_begin = (long long)blockIdx.x * w.rows_per_cta;
    const long long r_end = min(w.M, r_begin + w.rows_per_cta);
    const int nsteps = r_end > r_begin ? (int)((r_end - r_begin + ROWS - 1) / ROWS) : 0;
    const int ones_col = (w.db_part && blockIdx.y == gridDim.y - 1 && (w.Kin % BM) != 0) ? (w.Kin - k0) : -1;

    if (tid == 0) {
        for (int i = 0; i < WT_MAX_RAW; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_free[i], 2 * WT_GROUP / ARRIVE_DIV); }
        for (int i = 0; i < WT_MAX_TA; ++i) { mbar_init(&a_ready[i], WT_GROUP / ARRIVE_DIV); mbar_init(&a_free[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&b_ready[i], WT_GROUP / ARRIVE_DIV); mbar_init(&b_free[i], 1); }
        mbar_init(&all_done, 1);
        s_err = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    bool ok = true;
    if (warp < WT_A_THREADS / 32) {
        // ======================= A converters: x -> tensor memory =======================
        // warp (group g, quarter q): channels 32 q .. 32 q + 31 (= its TMEM lane quarter), ALL rows of the steps g, g + 2, ...
        const int q = warp & 3, g = warp >> 2;
        const int ch = q * 32 + lane;
        const bool has_data = ch < box_a;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ACC_COLS;
        int rs = g % D, ru = g / D, as = g % TA, au = g / TA;
        for (int it = g; it < nsteps; it += 2) {
            ok = mbar_wait(&raw_full[rs], (uint32_t)(ru & 1)) && ok;
            if ((warp & 3) == 0) WTL(0, it, 0);
            if (au >= 1) ok = mbar_wait(&a_free[as], (uint32_t)((au - 1) & 1)) && ok;   // MMAs that read the stage retired
            if ((warp & 3) == 0) WTL(0, it, 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float *src = reinterpret_cast<const float *>(raw_ring + (size_t)rs * raw_bytes) + ch;
            const uint32_t ta = t_lane + (uint32_t)(as * A_COLS);
#pragma unroll
            for (int c = 0; c < ROWS / 16; ++c) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = has_data ? src[(c * 16 + i) * box_a] : 0.f;
                if (ch == ones_col) {   // bias gradient: a row of ones for the valid rows
                    const long long row0 = r_begin + (long long)it * ROWS + c * 16;
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = row0 + i < w.M ? 1.f : 0.f;
                }
                if (split) {
                    float h[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) { h[i] = tf32_rn(v[i]); v[i] -= h[i]; }
                    tmem_st16(ta + c * 16, h);
                    tmem_st16(ta + ROWS + c * 16, v);
                } else {
                    tmem_st16(ta + c * 16, v);
                }
            }
            mbar_arrive_warp(&raw_free[rs]);   // the stores consumed every loaded value: the shared-memory reads have completed
            if ((warp & 3) == 0) WTL(0, it, 2);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive_warp(&a_ready[as]);
            if ((warp & 3) == 0) WTL(0, it, 3);
            rs += 2; if (rs >= D) { rs -= D; ++ru; }
            as += 2; if (as >= TA) { as -= TA; ++au; }
        }
    } else if (warp < (WT_A_THREADS + WT_B_THREADS) / 32) {
        // ======================= B converters: dy -> MN-major hi/lo images in shared memory =======================
        // group g owns operand stage g and the steps g, g + 2, ...
        const int bt = (tid - WT_A_THREADS) & (WT_GROUP - 1), g = (tid - WT_A_THREADS) / WT_GROUP;
        constexpr int CHUNKS = ROWS * BN / 4, CPT = CHUNKS / WT_GROUP;
        static_assert(CHUNKS % WT_GROUP == 0, "whole chunks per thread");
        uint32_t c_op[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int idx = bt + WT_GROUP * i;
            const int r = idx / (BN / 4), c4 = (idx % (BN / 4)) * 4;
            c_op[i] = (uint32_t)((c4 >> 5) * (ROWS * 128)) + sw128_32b(r, c4 & 31);
        }
        char *st_base = b_ring + (size_t)g * B_STAGE;
        int rs = g % D, ru = g / D, u = 0;
        for (int it = g; it < nsteps; it += 2, ++u) {
            ok = mbar_wait(&raw_full[rs], (uint32_t)(ru & 1)) && ok;
            if (bt < 32) WTL(1, it, 0);
            if (u >= 1) ok = mbar_wait(&b_free[g], (uint32_t)((u - 1) & 1)) && ok;
            if (bt < 32) WTL(1, it, 1);
            const char *src = raw_ring + (size_t)rs * raw_bytes + a_raw_bytes + (size_t)bt * 16;
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const float4 v = *reinterpret_cast<const float4 *>(src + i * (WT_GROUP * 16));
                char *dst = st_base + c_op[i];
                if (split) {
                    const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
                    *reinterpret_cast<float4 *>(dst) = h;
                    *reinterpret_cast<float4 *>(dst + B_BYTES) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                } else {
                    *reinterpret_cast<float4 *>(dst) = v;
                }
            }
            mbar_arrive_warp(&raw_free[rs]);
            if (bt < 32) WTL(1, it, 2);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_warp(&b_ready[g]);
            if (bt < 32) WTL(1, it, 3);
            rs += 2; if (rs >= D) { rs -= D; ++ru; }
        }
    } else if (warp == (WT_A_THREADS + WT_B_THREADS) / 32 + 1) {
        // ======================= loader: two TMA tile copies per step =======================
        const bool leader = elect_one();
        int rs = 0, ru = 0;
        for (int it = 0; it < nsteps; ++it) {
            if (ru >= 1) ok = mbar_wait(&raw_free[rs], (uint32_t)((ru - 1) & 1)) && ok;
            WTL(3, it, 0);
            if (leader) {
                char *slot = raw_ring + (size_t)rs * raw_bytes;
                const int row0 = (int)(r_begin + (long long)it * ROWS);
                mbar_expect_tx(&raw_full[rs], (uint32_t)raw_bytes);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(slot)), "l"(&tmap_x), "r"(k0), "r"(row0), "r"(smem_u32(&raw_full[rs])) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(slot + a_raw_bytes)), "l"(&tmap_g), "r"(n0), "r"(row0), "r"(smem_u32(&raw_full[rs])) : "memory");
            }
            if (++rs == D) { rs = 0; ++ru; }
        }
    } else {
        // ======================= MMA issuer: converged warp, one elected lane issues =======================
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_ts_mn(BN);
        const uint64_t bdesc0 = make_desc_mn(smem_u32(b_ring), ROWS * 128);
        int as = 0, au = 0;
        for (int it = 0; it < nsteps; ++it) {
            const int s = it & 1, u = it >> 1;
            ok = mbar_wait(&a_ready[as], (uint32_t)(au & 1)) && ok;
            WTL(2, it, 0);
            ok = mbar_wait(&b_ready[s], (uint32_t)(u & 1)) && ok;
            WTL(2, it, 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_tmem = tmem_base + (uint32_t)(ACC_COLS + as * A_COLS);
            const uint64_t dbh0 = bdesc0 + (uint64_t)((s * B_STAGE) >> 4), dbl0 = dbh0 + (uint64_t)(B_BYTES >> 4);
            if (leader) {
#pragma unroll
                for (int j = 0; j < ROWS / UMMA_K; ++j) {  // 8 rows = two 512-byte k groups per slab
                    const uint64_t o = (uint64_t)((j * 1024) >> 4);
                    umma_tf32_ts(tmem_base, a_tmem + j * UMMA_K, dbh0 + o, idesc, (it > 0 || j > 0) ? 1u : 0u);
                    if (split) {
                        umma_tf32_ts(tmem_base, a_tmem + j * UMMA_K, dbl0 + o, idesc, 1u);
                        umma_tf32_ts(tmem_base, a_tmem + ROWS + j * UMMA_K, dbh0 + o, idesc, 1u);
                    }
                }
                WTL(2, it, 2);
                umma_commit(&a_free[as]);
                umma_commit(&b_free[s]);
                if (it == nsteps - 1) umma_commit(&all_done);
            }
            WTL(2, it, 3);
            if (++as == TA) { as = 0; ++au; }
        }
    }
    if (warp < 8 && nsteps > 0) ok = mbar_wait(&all_done, 0u) && ok;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) s_err = 1;
    // epilogue: TMEM lane = Kin index (k0 + lane), column = N index; warps 0-3 and 4-7 split the columns
    if (warp < 8) {
        const int quarter = warp & 3, chalf = warp >> 2;
        const int m = quarter * 32 + lane;
        const int gk = k0 + m;
        float *prow = w.part + ((size_t)blockIdx.x * w.Kin + gk) * w.N + n0;
#pragma unroll
        for (int cc = 0; cc < BN / 2; cc += 16) {
            const int c0 = chalf * (BN / 2) + cc;
            float vals[16];
            if (nsteps > 0) tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, vals);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) vals[i] = 0.f;
            }
            if (gk < w.Kin) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (n0 + c0 + i < w.N) prow[c0 + i] = vals[i];
            } else if (m == ones_col && w.db_part) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (n0 + c0 + i < w.N) w.db_part[(size_t)blockIdx.x * w.N + n0 + c0 + i] = vals[i];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
    if (tid == 0 && s_err && w.error_flag) *w.error_flag = 1;
}

// un-swizzled 2-D tensor map over a row-major matrix [rows, cols] (row stride ld floats): box = box_rows x box_cols
static int make_tmap_plain(CUtensorMap *map, const float *base, long long rows, int cols, int ld, int box_cols, int box_rows) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return PU_ERR_UNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PU_OK : PU_ERR_INVALID_ARG;
}

static long long *g_wgrad_timeline = nullptr;   // development aid, see pu_tc_debug_set_wgrad_timeline
struct WtPlan { int box_a, rows, depth; size_t smem; };
static WtPlan wgrad_ts_plan(int Kin, int bn) {
    WtPlan pl{};
    const int cols = Kin < BM ? Kin : BM;   // widest A block any CTA sees
    pl.box_a = cols <= 32 ? 32 : (cols <= 64 ? 64 : 128);
    pl.rows = (pl.box_a <= 64 && bn <= 64) ? 64 : 32;   // narrow pairs: 64-row steps
    const size_t b_stage = 2 * (size_t)pl.rows * bn * 4, raw = (size_t)pl.rows * (pl.box_a + bn) * 4;
    long long d = (long long)((kMaxDynSmem - 2 * b_stage - 1024) / raw);
    pl.depth = (int)(d > WT_MAX_RAW ? WT_MAX_RAW : d);
    pl.smem = 2 * b_stage + (size_t)pl.depth * raw + 1024;
    return pl;
}

template <int BN, int ROWS>
static int launch_wgrad_ts_rows(const WParams &w0, const WgPlan &pl, const WtPlan &tp, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        PU_CUDA_TRY(cudaFuncSetAttribute(tc_wgrad_ts_kernel<BN, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
        configured = true;
    }
    CUtensorMap tx, tg;
    int rc = make_tmap_plain(&tx, w0.X, w0.M, w0.Kin, w0.ldx, tp.box_a, ROWS);
    if (rc != PU_OK) return rc;
    rc = make_tmap_plain(&tg, w0.G, w0.M, w0.N, w0.ldg, BN, ROWS);
    if (rc != PU_OK) return rc;
    WTParams w{};
    w.M = w0.M; w.Kin = w0.Kin; w.N = w0.N; w.rows_per_cta = w0.rows_per_cta; w.part = w0.part; w.db_part = w0.db_part;
    w.mode = w0.mode; w.raw_depth = tp.depth; w.box_a = tp.box_a; w.error_flag = w0.error_flag;
    w.timeline = g_wgrad_timeline;
    dim3 grid(pl.gx, pl.gy, pl.gz);
    tc_wgrad_ts_kernel<BN, ROWS><<<grid, WT_TOTAL, tp.smem, st>>>(w, tx, tg);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

template <int BN>
static int launch_wgrad_ts(const WParams &w0, const WgPlan &pl, cudaStream_t st) {
    const WtPlan tp = wgrad_ts_plan(w0.Kin, BN);
    if (tp.depth < 2) return PU_ERR_UNSUPPORTED;
    if constexpr (BN <= 64) {
        if (tp.rows == 64) return launch_wgrad_ts_rows<BN, 64>(w0, pl, tp, st);
    }
    return launch_wgrad_ts_rows<BN, 32>(w0, pl, tp, st);
}

}  // namespace tc
}  // namespace pu

using namespace pu;

extern "C" {

/* development aid (tools/wgrad_timeline.py): device buffer of roles x steps x events clock64 stamps, or NULL to switch off */
__attribute__((visibility("default"))) void pu_tc_debug_set_wgrad_timeline(long long *device_buffer) { tc::g_wgrad_timeline = device_buffer; }
__attribute__((visibility("default"))) int pu_tc_debug_wgrad_timeline_dims(int *roles, int *steps, int *events) {
    *roles = tc::WTL_ROLES; *steps = tc::WTL_STEPS; *events = tc::WTL_EVENTS;
    return 0;
}

/* 1 if (M,K,N, strides) can run on the tensor-core path */
int pu_tc_linear_supported(long long M, int K, int N, int ldx, int ldwt, int ldy) {
    return M > 0 && K >= 32 && N >= 32 && (K & 3) == 0 && (ldx & 3) == 0 && (ldwt & 3) == 0 && (ldy >= N) && (N & 3) == 0;
}

/* scratch for the packed (pre-split, pre-swizzled) weight image of the streamed-weight kernels */
size_t pu_tc_workspace_bytes(int K, int N) { return tc::tc_workspace_bytes(K, N); }

int pu_tc_linear_fwd(const float *x, int ldx, const float *wt, int ldwt, const float *bias, float *y, int ldy, long long M,
                     int K, int N, int accumulate, float *stat_sum, float *stat_m2, int mode, int *error_flag,
                     void *workspace, size_t workspace_bytes, pu_stream_t stream) {
    return pu_tc_linear_fwd_ex(x, ldx, wt, ldwt, bias, y, ldy, M, K, N, accumulate, stat_sum, stat_m2, mode, error_flag, workspace,
                               workspace_bytes, PU_F32, stream);
}

int pu_tc_linear_fwd_ex(const float *x, int ldx, const float *wt, int ldwt, const float *bias, void *yv, int ldy, long long M,
                        int K, int N, int accumulate, float *stat_sum, float *stat_m2, int mode, int *error_flag,
                        void *workspace, size_t workspace_bytes, int y_dtype, pu_stream_t stream) {
    float *y = (float *)yv;
    if (y_dtype != PU_F32 && (y_dtype != PU_BF16 || accumulate)) return PU_ERR_INVALID_ARG;
    if (!x || !wt || !y || M < 0 || K < 1 || N < 1 || ldx < K || ldwt < K || ldy < N) return PU_ERR_INVALID_ARG;
    if ((stat_sum == nullptr) != (stat_m2 == nullptr)) return PU_ERR_INVALID_ARG;
    if (mode != 1 && mode != 3) return PU_ERR_INVALID_ARG;
    if (M == 0) return PU_OK;
    if (!pu_tc_linear_supported(M, K, N, ldx, ldwt, ldy) || ((((uintptr_t)x) | ((uintptr_t)wt)) & 15)) return PU_ERR_UNSUPPORTED;
    tc::Params2 q{};
    q.g.A = x; q.g.lda = ldx; q.g.Bt = wt; q.g.ldb = ldwt; q.g.C = y; q.g.ldc = ldy; q.g.bias = bias;
    q.g.M = M; q.g.N = N; q.g.K = K; q.g.accumulate = accumulate; q.g.stat_sum = stat_sum; q.g.stat_m2 = stat_m2; q.g.mode = mode;
    q.g.error_flag = error_flag;
    q.g.c_bf16 = y_dtype == PU_BF16;
    q.ntiles = (M + tc::BM - 1) / tc::BM;
    return tc::dispatch_persist<tc::EPI_STORE>(q, workspace, workspace_bytes, (cudaStream_t)stream);
}

/* 1 if the fused att_pooling kernels can run on the tensor-core path for channel width d */
int pu_tc_att_supported(int K, int d, int ldx) { return K == 16 && d >= 32 && (d & 3) == 0 && (ldx & 3) == 0; }

int pu_tc_att_pooling_fwd(const float *feature_set, int ldx, const float *wt, long long P, int K, int d, float *f_agg,
                          int ldo, int mode, int *error_flag, void *workspace, size_t workspace_bytes, pu_stream_t stream) {
    if (!feature_set || !wt || !f_agg || P < 0 || ldx < d || ldo < d) return PU_ERR_INVALID_ARG;
    if (mode != 1 && mode != 3) return PU_ERR_INVALID_ARG;
    if (!pu_tc_att_supported(K, d, ldx) || (((uintptr_t)feature_set | (uintptr_t)wt) & 15)) return PU_ERR_UNSUPPORTED;
    if (P == 0) return PU_OK;
    tc::Params2 q{};
    q.g.A = feature_set; q.g.lda = ldx; q.g.Bt = wt; q.g.ldb = d; q.g.M = P * K; q.g.N = d; q.g.K = d; q.g.mode = mode;
    q.g.error_flag = error_flag;
    q.X = feature_set; q.ldx = ldx; q.OUT = f_agg; q.ldo = ldo;
    q.ntiles = (q.g.M + tc::BM - 1) / tc::BM;
    return tc::dispatch_persist<tc::EPI_ATT_FWD>(q, workspace, workspace_bytes, (cudaStream_t)stream);
}

int pu_tc_att_pooling_bwd(const float *feature_set, int ldx, const float *wt, const float *g_agg, int ldg, long long P,
                          int K, int d, float *d_act, int ldda, float *dx_direct, int lddx, int mode, int *error_flag,
                          void *workspace, size_t workspace_bytes, pu_stream_t stream) {
    if (!feature_set || !wt || !g_agg || !d_act || !dx_direct || P < 0 || ldx < d || ldg < d || ldda < d || lddx < d)
        return PU_ERR_INVALID_ARG;
    if (mode != 1 && mode != 3) return PU_ERR_INVALID_ARG;
    if (!pu_tc_att_supported(K, d, ldx) || (((uintptr_t)feature_set | (uintptr_t)wt) & 15)) return PU_ERR_UNSUPPORTED;
    if (P == 0) return PU_OK;
    tc::Params2 q{};
    q.g.A = feature_set; q.g.lda = ldx; q.g.Bt = wt; q.g.ldb = d; q.g.M = P * K; q.g.N = d; q.g.K = d; q.g.mode = mode;
    q.g.C = d_act; q.g.ldc = ldda; q.g.error_flag = error_flag;
    q.X = feature_set; q.ldx = ldx; q.G = g_agg; q.ldg = ldg; q.OUT = dx_direct; q.ldo = lddx;
    q.ntiles = (q.g.M + tc::BM - 1) / tc::BM;
    return tc::dispatch_persist<tc::EPI_ATT_BWD>(q, workspace, workspace_bytes, (cudaStream_t)stream);
}

/* Fused att_pooling backward for d = 64 (K = 16): as pu_tc_att_pooling_bwd, but `dx` receives the COMPLETE gradient
 * g s + d_act w^T (the dgrad through the FC runs as a second tensor-core MMA inside the kernel); d_act is still written for
 * the weight gradient.  `w` is the FC kernel [d, d] in its own orientation, `wt` its transpose. */
int pu_tc_att_bwd_fused_supported(int K, int d, int ldx) { return K == 16 && d == 64 && (ldx & 3) == 0; }

int pu_tc_att_pooling_bwd_fused(const float *feature_set, int ldx, const float *wt, const float *w, const float *g_agg, int ldg,
                                long long P, int K, int d, float *d_act, int ldda, float *dx, int lddx, int mode,
                                int *error_flag, pu_stream_t stream) {
    if (!feature_set || !wt || !w || !g_agg || !d_act || !dx || P < 0 || ldx < d || ldg < d || ldda < d || lddx < d)
        return PU_ERR_INVALID_ARG;
    if (mode != 1 && mode != 3) return PU_ERR_INVALID_ARG;
    if (!pu_tc_att_bwd_fused_supported(K, d, ldx) || (lddx & 3) ||
        (((uintptr_t)feature_set | (uintptr_t)wt | (uintptr_t)w | (uintptr_t)dx) & 15))
        return PU_ERR_UNSUPPORTED;
    if (P == 0) return PU_OK;
    tc::Params2 q{};
    q.g.A = feature_set; q.g.lda = ldx; q.g.Bt = wt; q.g.ldb = d; q.g.M = P * K; q.g.N = d; q.g.K = d; q.g.mode = mode;
    q.g.C = d_act; q.g.ldc = ldda; q.g.error_flag = error_flag;
    q.X = feature_set; q.ldx = ldx; q.G = g_agg; q.ldg = ldg; q.OUT = dx; q.ldo = lddx; q.W2 = w; q.ldw2 = d;
    q.ntiles = (q.g.M + tc::BM - 1) / tc::BM;
    return tc::launch_persist<64, tc::EPI_ATT_BWD_F, false>(q, nullptr, 0, (cudaStream_t)stream);
}

/* Tensor-core weight gradient: dw[Kin,N] (+)= x^T dy, db[N] (+)= column sums of dy (db only when Kin % 128 != 0).
 * Supported when Kin >= 32 or N > 32, all of Kin, N, ldx, lddy multiples of 4. */
int pu_tc_wgrad_supported(long long M, int Kin, int N, int ldx, int lddy, int want_db) {
    if (M < 512 || (Kin & 3) || (N & 3) || (ldx & 3) || (lddy & 3) || N < 32 || Kin < 32) return 0;
    if (want_db && (Kin % tc::BM) == 0) return 0;
    return 1;
}

size_t pu_tc_wgrad_workspace_bytes(long long M, int Kin, int N) {
    const tc::WgPlan pl = tc::wgrad_plan_tc(M, Kin, N);
    return ((size_t)pl.gx * Kin * N + (size_t)pl.gx * N) * sizeof(float) + 256;
}

int pu_tc_wgrad(const float *x, int ldx, const float *dy, int lddy, long long M, int Kin, int N, float *dw, float *db,
                int accumulate, int mode, void *workspace, size_t workspace_bytes, int *error_flag, pu_stream_t stream) {
    if (!x || !dy || !dw || M < 1 || Kin < 1 || N < 1 || ldx < Kin || lddy < N) return PU_ERR_INVALID_ARG;
    if (mode != 1 && mode != 3) return PU_ERR_INVALID_ARG;
    if (!pu_tc_wgrad_supported(M, Kin, N, ldx, lddy, db != nullptr) || ((((uintptr_t)x) | ((uintptr_t)dy)) & 15))
        return PU_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < pu_tc_wgrad_workspace_bytes(M, Kin, N)) return PU_ERR_WORKSPACE;
    const tc::WgPlan pl = tc::wgrad_plan_tc(M, Kin, N);
    if (pl.depth < 2) return PU_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    tc::WParams w{};
    w.X = x; w.ldx = ldx; w.G = dy; w.ldg = lddy; w.M = M; w.Kin = Kin; w.N = N; w.rows_per_cta = pl.rows_per_cta;
    w.part = (float *)workspace;
    w.db_part = db ? w.part + (size_t)pl.gx * Kin * N : nullptr;
    w.mode = mode; w.raw_depth = pl.depth; w.error_flag = error_flag;
    // PU_WGRAD_TS=0 selects the first-generation kernel (both operands in shared memory) for A/B runs
    static const bool use_ts = [] { const char *e = getenv("PU_WGRAD_TS"); return !(e && e[0] == '0'); }();
    int rc;
    if (use_ts) {
        if (pl.bn == 32) rc = tc::launch_wgrad_ts<32>(w, pl, st);
        else if (pl.bn == 64) rc = tc::launch_wgrad_ts<64>(w, pl, st);
        else rc = tc::launch_wgrad_ts<128>(w, pl, st);
    } else {
        if (pl.bn == 32) rc = tc::launch_wgrad<32>(w, pl, st);
        else if (pl.bn == 64) rc = tc::launch_wgrad<64>(w, pl, st);
        else rc = tc::launch_wgrad<128>(w, pl, st);
    }
    if (rc != PU_OK) return rc;
    launch_reduce_parts(w.part, pl.gx, (long long)Kin * N, dw, accumulate, st);
    PU_LAUNCH_CHECK();
    if (db) {
        launch_reduce_parts(w.db_part, pl.gx, N, db, accumulate, st);
        PU_LAUNCH_CHECK();
    }
    return PU_OK;
}

}  // extern "C"
