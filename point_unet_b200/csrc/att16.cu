// att16.cu -- Network.att_pooling (PointSegment/RandLANet.py:388-401, up to f_agg) for the 16-channel level of the
// encoder (d = 16, K = 16: Encoder_layer_0 of PointSegment, 4 x 180 000 points per step), forward and FULLY FUSED
// backward on CUDA cores.
//
// Why a kernel of its own.  At d = 16 the op is a 16x16x16 product per point: far too small for a tcgen05 tile (the
// tensor-core kernels of tc_gemm.cu start at d = 32) and HBM-bound by a wide margin (1 KB of x per point).  The generic
// tiled kernels of mlp.cu needed three passes for the backward -- d_act and g*s written to HBM, then dx += d_act w^T and
// dw = x^T d_act read them back: 8 round trips over the [P,16,16] tensor.  Here one pass reads x once and writes dx once:
//
//   act = x w            (16 rows x 16 x 16)          s = softmax over the 16 neighbours, per channel
//   f_agg = sum_k x s                                                                  (forward)
//   d_act = s (g x - sum_k g x s)      dx = g s + d_act w^T      dw += x^T d_act       (backward)
//
// Mapping.  16 lanes own one point, a warp owns two.  A lane plays two roles: ROW owner (neighbour k = its x row in 16
// registers) for the two products against w, and COLUMN owner (channel c) for the softmax over K, which is then
// thread-local.  The FC kernel sits in CONSTANT memory: both products are FFMAs whose second operand is a uniform register
// filled by LDCU.128 on the uniform datapath -- no shared-memory traffic for w at all (a broadcast LDS.128 per 4 FMAs
// would make the kernel shared-memory bound).  Role changes go through padded shared-memory tiles (row stride 20 floats, the two points of a warp 16 banks
// apart: every access below is conflict-free).  dw is accumulated as a 4x4 register block per lane over all the points
// a lane sees and reduced once per CTA in a fixed order (deterministic).  x tiles arrive through a per-warp cp.async
// ring (3 stages, 2 KB per stage), dx leaves through the same tile with full-line 128-bit stores.
#include "common.cuh"

namespace pu {
namespace att16 {

constexpr int D = 16;              // channels
constexpr int KN = 16;             // neighbours
constexpr int RS = 20;             // padded row stride of a tile (floats)
constexpr int TILE = KN * RS + 16; // floats per point tile; 336 = 16 (mod 32): the two half-warps use disjoint banks
constexpr int WARPS = 8;
constexpr int STAGES = 3;
constexpr int SLOTS = 4;           // constant-memory copies of w in flight (one per call, round robin)
constexpr int FWD_TILES = STAGES * 2 + 2;      // x ring + T
constexpr int BWD_TILES = STAGES * 2 + 4;      // x ring + T/E + D
constexpr float LOG2E = 1.4426950408889634f;

}  // namespace att16
}  // namespace pu
// w[slot][j][c] (tf.layers.dense kernel [in, out]); C linkage so that the inline PTX below can name it
extern "C" { __constant__ float pu_att16_cw[4 * 16 * 16]; __constant__ float pu_att16_cw2[4 * 16 * 16]; }
namespace pu {
namespace att16 {

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void load_row16(const float *p, float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 a = *reinterpret_cast<const float4 *>(p + 4 * q);
        v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
    }
}
__device__ __forceinline__ void store_row16(float *p, const float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4 *>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// stage the x tiles of point pair `pair` (2 KB, coalesced 16-byte chunks); points past the end are zero-filled
// The 16 channels of a row may live in two tensors (xa: channels 0-7, row stride lda; xb: channels 8-15, row stride ldb) --
// the two halves of building_block's concat kept as separate, contiguous tensors (a 32-byte half row inside a 64-byte row
// costs a full DRAM burst per half).  One concat tensor is the special case xb = xa + 8, ldb = lda.
// (a lane always moves the same 16-byte column chunk q = lane & 3 of a row, so the tensor and stride it addresses are fixed
// for its lifetime: xq = base of its chunk's tensor + column offset, ldq = that tensor's row stride)
__device__ __forceinline__ void issue_pair(const float *__restrict__ xq, int ldq, long long P, long long pair, float *stage,
                                           int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int chunk = lane + 32 * i, h = chunk >> 6, r = chunk & 63, k = r >> 2, q = r & 3;
        const long long p = pair * 2 + h;
        float *dst = stage + h * TILE + k * RS + q * 4;
        if (p < P) cp_async16(dst, xq + ((size_t)p * KN + k) * ldq);
        else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Packed fp32 pairs: Blackwell's FFMA2 (fma.rn.f32x2) does two FMAs per issue slot.  These kernels are issue-bound (ncu:
// 80-85 % of the issue slots busy, FMA pipe at 55 %), so halving the FMA instruction count is what makes them faster.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// Four consecutive constants as two packed pairs.  `volatile`: the load stays where it is written.  Left to itself the
// compiler treats the 256 weights as loop-invariant, hoists every load out of the point loop and spills them (1.8 KB of
// local-memory traffic per iteration); kept in the loop they become uniform-datapath loads (LDCU.128) and the FFMA2s take
// their second operand from uniform registers.  The backward reads the weights twice per iteration and uses a second copy
// of them (pu_att16_cw2) for the second product: with two uses of the same constant ptxas goes back to hoisting.
template <int OFF>  // OFF: float index into pu_att16_cw
__device__ __forceinline__ void ldc4(u64 &p01, u64 &p23) {
    asm volatile("ld.const.v2.b64 {%0, %1}, [pu_att16_cw+%2];" : "=l"(p01), "=l"(p23) : "n"(OFF * 4));
}
template <int OFF>
__device__ __forceinline__ void ldc4b(u64 &p01, u64 &p23) {
    asm volatile("ld.const.v2.b64 {%0, %1}, [pu_att16_cw2+%2];" : "=l"(p01), "=l"(p23) : "n"(OFF * 4));
}

// act[c] += xr[j] w[j][c]   (compile-time recursion over j so that every constant offset is an immediate);
// act2[i] = (act[2i], act[2i+1])
template <int SLOT, int J>
__device__ __forceinline__ void rtw_step(const float (&xr)[16], u64 (&act2)[8]) {
    u64 w[8];
    ldc4<SLOT * D * D + J * D + 0>(w[0], w[1]);
    ldc4<SLOT * D * D + J * D + 4>(w[2], w[3]);
    ldc4<SLOT * D * D + J * D + 8>(w[4], w[5]);
    ldc4<SLOT * D * D + J * D + 12>(w[6], w[7]);
    const u64 xx = pack2(xr[J], xr[J]);
#pragma unroll
    for (int i = 0; i < 8; ++i) act2[i] = fma2(xx, w[i], act2[i]);
    if constexpr (J + 1 < D) rtw_step<SLOT, J + 1>(xr, act2);
}
template <int SLOT>
__device__ __forceinline__ void row_times_w(const float (&xr)[16], float (&act)[16]) {
    u64 act2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) act2[i] = 0ull;
    rtw_step<SLOT, 0>(xr, act2);
#pragma unroll
    for (int i = 0; i < 8; ++i) unpack2(act2[i], act[2 * i], act[2 * i + 1]);
}
// o[j] += sum_c dr[c] w[j][c]: pairs run over c, (even, odd) partial sums per output channel
template <int SLOT, int J>
__device__ __forceinline__ void rtwt_step(const u64 (&dr2)[8], float (&o)[16]) {
    u64 w[8];
    ldc4b<SLOT * D * D + J * D + 0>(w[0], w[1]);
    ldc4b<SLOT * D * D + J * D + 4>(w[2], w[3]);
    ldc4b<SLOT * D * D + J * D + 8>(w[4], w[5]);
    ldc4b<SLOT * D * D + J * D + 12>(w[6], w[7]);
    u64 s0 = pack2(o[J], 0.f), s1 = 0ull;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        s0 = fma2(dr2[i], w[i], s0);
        s1 = fma2(dr2[i + 1], w[i + 1], s1);
    }
    float a, b, c, d;
    unpack2(s0, a, b);
    unpack2(s1, c, d);
    o[J] = (a + b) + (c + d);
    if constexpr (J + 1 < D) rtwt_step<SLOT, J + 1>(dr2, o);
}
template <int SLOT>
__device__ __forceinline__ void row_times_wt(const float (&dr)[16], float (&o)[16]) {
    u64 dr2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dr2[i] = pack2(dr[2 * i], dr[2 * i + 1]);
    rtwt_step<SLOT, 0>(dr2, o);
}

// softmax over the 16 values of a column (in place: a[k] <- e_k, returns 1 / sum)
__device__ __forceinline__ float softmax16(float (&a)[16]) {
    float m = a[0];
#pragma unroll
    for (int k = 1; k < KN; ++k) m = fmaxf(m, a[k]);
    const float mb = -m * LOG2E;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < KN; ++k) {
        a[k] = ex2_approx(fmaf(a[k], LOG2E, mb));
        sum += a[k];
    }
    return rcp_approx(sum);
}

template <int SLOT>
__global__ void __launch_bounds__(WARPS * 32, 2)
    att16_fwd_kernel(const float *__restrict__ x, int ldx, const float *__restrict__ xb, int ldxb, long long P,
                     float *__restrict__ out, int ldo) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, half = lane >> 4, t = lane & 15;
    const float *xq = (lane & 2) ? xb + ((lane & 3) - 2) * 4 : x + (lane & 3) * 4;
    const int ldq = (lane & 2) ? ldxb : ldx;
    float *wb = smem + (size_t)wib * FWD_TILES * TILE;
    float *T = wb + (STAGES * 2 + half) * TILE;
    const long long npairs = (P + 1) >> 1;
    const long long nw = (long long)gridDim.x * WARPS;
    long long pair = (long long)blockIdx.x * WARPS + wib;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (pair + s * nw < npairs) issue_pair(xq, ldq, P, pair + s * nw, wb + s * 2 * TILE, lane);
        cp_async_commit();
    }
#pragma unroll 1
    for (int it = 0; pair < npairs; pair += nw, ++it) {
        const int stage = it % STAGES;
        {
            const long long nxt = pair + (STAGES - 1) * nw;
            if (nxt < npairs) issue_pair(xq, ldq, P, nxt, wb + ((it + STAGES - 1) % STAGES) * 2 * TILE, lane);
            cp_async_commit();
        }
        cp_async_wait<STAGES - 1>();
        __syncwarp();
        const float *X = wb + (stage * 2 + half) * TILE;
        const long long p = pair * 2 + half;

        float a[16];
        {
            float xr[16];
            load_row16(X + t * RS, xr);       // row owner: neighbour k = t
            row_times_w<SLOT>(xr, a);
        }
#pragma unroll
        for (int c = 0; c < D; ++c) T[c * RS + t] = a[c];   // transposed: T[c][k]
        __syncwarp();
        load_row16(T + t * RS, a);            // column owner: channel c = t, a[k] = act[k][c]
        const float inv = softmax16(a);
        float num = 0.f;
#pragma unroll
        for (int k = 0; k < KN; ++k) num = fmaf(X[k * RS + t], a[k], num);
        if (p < P) out[(size_t)p * ldo + t] = num * inv;
        __syncwarp();  // the tiles of this stage and T are free again
    }
}

template <int SLOT>
__global__ void __launch_bounds__(WARPS * 32, 2)
    att16_bwd_kernel(const float *__restrict__ x, int ldx, const float *__restrict__ xb, int ldxb,
                     const float *__restrict__ g_agg, int ldg, long long P, float *__restrict__ dx, int lddx,
                     float *__restrict__ dxb, int lddxb, float *__restrict__ dw_part) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, half = lane >> 4, t = lane & 15;
    const int jq = t >> 2, cq = t & 3;
    const float *xq = (lane & 2) ? xb + ((lane & 3) - 2) * 4 : x + (lane & 3) * 4;
    const int ldq = (lane & 2) ? ldxb : ldx;
    float *dxq = (lane & 2) ? dxb + ((lane & 3) - 2) * 4 : dx + (lane & 3) * 4;
    const int lddq = (lane & 2) ? lddxb : lddx;
    float *wb = smem + (size_t)wib * BWD_TILES * TILE;
    float *T = wb + (STAGES * 2 + half) * TILE;      // act^T, later E = g*s (row layout)
    float *Dt = wb + (STAGES * 2 + 2 + half) * TILE; // d_act (row layout)
    const long long npairs = (P + 1) >> 1;
    const long long nw = (long long)gridDim.x * WARPS;
    long long pair = (long long)blockIdx.x * WARPS + wib;

    u64 wacc[4][2];  // dw[4 jq + a][4 cq + (0,1)], [4 cq + (2,3)]
#pragma unroll
    for (int a = 0; a < 4; ++a) wacc[a][0] = wacc[a][1] = 0ull;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (pair + s * nw < npairs) issue_pair(xq, ldq, P, pair + s * nw, wb + s * 2 * TILE, lane);
        cp_async_commit();
    }
    float g_next = (pair * 2 + half < P) ? g_agg[(size_t)(pair * 2 + half) * ldg + t] : 0.f;
#pragma unroll 1
    for (int it = 0; pair < npairs; pair += nw, ++it) {
        const int stage = it % STAGES;
        {
            const long long nxt = pair + (STAGES - 1) * nw;
            if (nxt < npairs) issue_pair(xq, ldq, P, nxt, wb + ((it + STAGES - 1) % STAGES) * 2 * TILE, lane);
            cp_async_commit();
        }
        cp_async_wait<STAGES - 1>();
        __syncwarp();
        float *Xw = wb + stage * 2 * TILE;   // both points of the pair (copy-out)
        float *X = Xw + half * TILE;
        const float g = g_next;
        {
            const long long pn = (pair + nw) * 2 + half;  // upstream gradient of the next pair: in flight during this one
            g_next = (pn < P) ? g_agg[(size_t)pn * ldg + t] : 0.f;
        }

        float a[16];
        {
            float xr[16];
            load_row16(X + t * RS, xr);
            row_times_w<SLOT>(xr, a);
        }
#pragma unroll
        for (int c = 0; c < D; ++c) T[c * RS + t] = a[c];
        __syncwarp();
        load_row16(T + t * RS, a);            // column owner: a[k] = act[k][c = t]
        const float inv = softmax16(a);
        float dot = 0.f;
        float ds[16];
#pragma unroll
        for (int k = 0; k < KN; ++k) {
            a[k] *= inv;                       // s_k
            ds[k] = g * X[k * RS + t];         // d s_k = g x_k
            dot = fmaf(a[k], ds[k], dot);
        }
        __syncwarp();                          // every lane has read its T row: T becomes E
#pragma unroll
        for (int k = 0; k < KN; ++k) {
            Dt[k * RS + t] = a[k] * (ds[k] - dot);   // d_act[k][c]
            T[k * RS + t] = g * a[k];                // direct term g s
        }
        __syncwarp();
        // dw[4jq.., 4cq..] += sum_k x[k][4jq..] (x) d_act[k][4cq..]
#pragma unroll
        for (int k = 0; k < KN; ++k) {
            const float4 xv = *reinterpret_cast<const float4 *>(X + k * RS + 4 * jq);
            const float4 dv = *reinterpret_cast<const float4 *>(Dt + k * RS + 4 * cq);
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
            const u64 d01 = pack2(dv.x, dv.y), d23 = pack2(dv.z, dv.w);
#pragma unroll
            for (int a2 = 0; a2 < 4; ++a2) {
                const u64 xx = pack2(xa[a2], xa[a2]);
                wacc[a2][0] = fma2(xx, d01, wacc[a2][0]);
                wacc[a2][1] = fma2(xx, d23, wacc[a2][1]);
            }
        }
        // row owner again: dx[k = t][j] = g s + sum_c d_act[k][c] w[j][c]
        float o[16];
        {
            float dr[16];
            load_row16(Dt + t * RS, dr);
            load_row16(T + t * RS, o);
            row_times_wt<SLOT>(dr, o);
        }
        __syncwarp();                          // all reads of X (dw product) are done: the tile now carries dx
        store_row16(X + t * RS, o);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int chunk = lane + 32 * i, h = chunk >> 6, r = chunk & 63, k = r >> 2, q = r & 3;
            const long long pp = pair * 2 + h;
            if (pp < P)
                st_stream_f4(reinterpret_cast<float4 *>(dxq + ((size_t)pp * KN + k) * lddq),
                             *reinterpret_cast<const float4 *>(Xw + h * TILE + k * RS + q * 4));
        }
        __syncwarp();
    }

    // per-CTA partial of dw: 16 (warp, half) groups summed in a fixed order
    cp_async_wait<0>();
    __syncthreads();
    float *red = smem;  // [16 groups][256]
    {
        float *mine = red + (wib * 2 + half) * (D * D);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float4 v;
            unpack2(wacc[a][0], v.x, v.y);
            unpack2(wacc[a][1], v.z, v.w);
            *reinterpret_cast<float4 *>(mine + (4 * jq + a) * D + 4 * cq) = v;
        }
    }
    __syncthreads();
    {
        float s = 0.f;
#pragma unroll
        for (int gI = 0; gI < WARPS * 2; ++gI) s += red[gI * (D * D) + threadIdx.x];
        dw_part[(size_t)blockIdx.x * (D * D) + threadIdx.x] = s;
    }
}

static int plan_grid(long long P) {
    const long long npairs = (P + 1) / 2;
    long long ctas = (npairs + WARPS - 1) / WARPS;
    const long long cap = 2LL * kNumSMs;  // two resident CTAs per SM, each warp walks its pairs with a grid stride
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    return (int)ctas;
}

// The FC kernel travels through one of SLOTS __constant__ copies.  Work queued on ONE stream is serialised, so a stream may
// keep reusing its slot; two streams must never share one (the second copy would overwrite weights a running kernel still
// reads).  Slots are therefore handed out per (device, stream) and recycled least-recently-used: a race would need more
// than SLOTS streams with att16 kernels in flight at the same time (the training step uses one).
struct SlotOwner { int dev; cudaStream_t st; unsigned long long last; bool used; };
static SlotOwner g_owner[SLOTS] = {};
static unsigned long long g_tick = 0;
static int g_slot_lock = 0;
static int pick_slot(cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    while (__atomic_exchange_n(&g_slot_lock, 1, __ATOMIC_ACQUIRE)) { }
    int mine = -1, empty = -1, lru = 0;
    for (int i = 0; i < SLOTS; ++i) {
        if (!g_owner[i].used) { if (empty < 0) empty = i; continue; }
        if (g_owner[i].dev == dev && g_owner[i].st == st) { mine = i; break; }
        if (!g_owner[lru].used || g_owner[i].last < g_owner[lru].last) lru = i;
    }
    // own slot, else a free one, else recycle the least recently used (all SLOTS taken by other streams)
    const int pick = mine >= 0 ? mine : (empty >= 0 ? empty : lru);
    g_owner[pick] = SlotOwner{dev, st, ++g_tick, true};
    __atomic_store_n(&g_slot_lock, 0, __ATOMIC_RELEASE);
    return pick;
}

template <int SLOT>
static int launch_fwd(const float *x, int ldx, const float *xb, int ldxb, const float *w, long long P, float *out, int ldo,
                      cudaStream_t st) {
    constexpr size_t smem = (size_t)WARPS * FWD_TILES * TILE * sizeof(float);
    static bool attr[64] = {};  // per device: the attribute belongs to the device's copy of the function
    int dev = 0;
    PU_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        PU_CUDA_TRY(cudaFuncSetAttribute(att16_fwd_kernel<SLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    PU_CUDA_TRY(cudaMemcpyToSymbolAsync(pu_att16_cw, w, D * D * sizeof(float), (size_t)SLOT * D * D * sizeof(float),
                                        cudaMemcpyDeviceToDevice, st));
    att16_fwd_kernel<SLOT><<<plan_grid(P), WARPS * 32, smem, st>>>(x, ldx, xb, ldxb, P, out, ldo);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

template <int SLOT>
static int launch_bwd(const float *x, int ldx, const float *xb, int ldxb, const float *w, const float *g, int ldg, long long P,
                      float *dx, int lddx, float *dxb, int lddxb, float *dw, int accumulate, float *part, cudaStream_t st) {
    constexpr size_t smem = (size_t)WARPS * BWD_TILES * TILE * sizeof(float);
    static_assert(WARPS * BWD_TILES * TILE >= WARPS * 2 * D * D, "reduction scratch fits");
    static bool attr[64] = {};
    int dev = 0;
    PU_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        PU_CUDA_TRY(cudaFuncSetAttribute(att16_bwd_kernel<SLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    PU_CUDA_TRY(cudaMemcpyToSymbolAsync(pu_att16_cw, w, D * D * sizeof(float), (size_t)SLOT * D * D * sizeof(float),
                                        cudaMemcpyDeviceToDevice, st));
    PU_CUDA_TRY(cudaMemcpyToSymbolAsync(pu_att16_cw2, w, D * D * sizeof(float), (size_t)SLOT * D * D * sizeof(float),
                                        cudaMemcpyDeviceToDevice, st));
    const int grid = plan_grid(P);
    att16_bwd_kernel<SLOT><<<grid, WARPS * 32, smem, st>>>(x, ldx, xb, ldxb, g, ldg, P, dx, lddx, dxb, lddxb, part);
    PU_LAUNCH_CHECK();
    launch_reduce_parts(part, grid, D * D, dw, accumulate, st);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

}  // namespace att16
}  // namespace pu

extern "C" {

using namespace pu;
using namespace pu::att16;

int pu_att16_supported(int K, int d, int ldx) { return (K == KN && d == D && ldx >= D && (ldx & 3) == 0) ? 1 : 0; }

size_t pu_att16_workspace_bytes(long long P) {
    if (P < 0) return 0;
    return (size_t)plan_grid(P) * D * D * sizeof(float) + 256;
}

int pu_att16_supported_split(int K, int d, int ld_lo, int ld_hi) {
    return (K == KN && d == D && ld_lo >= D / 2 && ld_hi >= D / 2 && ((ld_lo | ld_hi) & 3) == 0) ? 1 : 0;
}

int pu_att16_fwd_split(const float *x_lo, int ld_lo, const float *x_hi, int ld_hi, const float *w, long long P, float *f_agg,
                       int ldo, pu_stream_t stream) {
    if (!x_lo || !x_hi || !w || !f_agg || P < 0 || ld_lo < D / 2 || ld_hi < D / 2 || ((ld_lo | ld_hi) & 3) || ldo < D ||
        ((((uintptr_t)x_lo) | ((uintptr_t)x_hi)) & 15))
        return PU_ERR_INVALID_ARG;
    if (P == 0) return PU_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pick_slot(st)) {
        case 0: return launch_fwd<0>(x_lo, ld_lo, x_hi, ld_hi, w, P, f_agg, ldo, st);
        case 1: return launch_fwd<1>(x_lo, ld_lo, x_hi, ld_hi, w, P, f_agg, ldo, st);
        case 2: return launch_fwd<2>(x_lo, ld_lo, x_hi, ld_hi, w, P, f_agg, ldo, st);
        default: return launch_fwd<3>(x_lo, ld_lo, x_hi, ld_hi, w, P, f_agg, ldo, st);
    }
}

int pu_att16_fwd(const float *feature_set, int ldx, const float *w, long long P, float *f_agg, int ldo,
                 pu_stream_t stream) {
    if (!feature_set || ldx < D) return PU_ERR_INVALID_ARG;
    return pu_att16_fwd_split(feature_set, ldx, feature_set + D / 2, ldx, w, P, f_agg, ldo, stream);
}

int pu_att16_bwd_split(const float *x_lo, int ld_lo, const float *x_hi, int ld_hi, const float *w, const float *g_agg, int ldg,
                       long long P, float *dx_lo, int lddx_lo, float *dx_hi, int lddx_hi, float *dw, int accumulate,
                       void *workspace, size_t workspace_bytes, pu_stream_t stream) {
    if (!x_lo || !x_hi || !w || !g_agg || !dx_lo || !dx_hi || !dw || P < 0 || ld_lo < D / 2 || ld_hi < D / 2 || ldg < D ||
        lddx_lo < D / 2 || lddx_hi < D / 2 || ((ld_lo | ld_hi | lddx_lo | lddx_hi) & 3) ||
        ((((uintptr_t)x_lo) | ((uintptr_t)x_hi) | ((uintptr_t)dx_lo) | ((uintptr_t)dx_hi)) & 15))
        return PU_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (P == 0) {
        if (!accumulate) PU_CUDA_TRY(cudaMemsetAsync(dw, 0, D * D * sizeof(float), st));
        return PU_OK;
    }
    if (!workspace || workspace_bytes < pu_att16_workspace_bytes(P)) return PU_ERR_WORKSPACE;
    float *part = (float *)workspace;
    switch (pick_slot(st)) {
        case 0: return launch_bwd<0>(x_lo, ld_lo, x_hi, ld_hi, w, g_agg, ldg, P, dx_lo, lddx_lo, dx_hi, lddx_hi, dw, accumulate, part, st);
        case 1: return launch_bwd<1>(x_lo, ld_lo, x_hi, ld_hi, w, g_agg, ldg, P, dx_lo, lddx_lo, dx_hi, lddx_hi, dw, accumulate, part, st);
        case 2: return launch_bwd<2>(x_lo, ld_lo, x_hi, ld_hi, w, g_agg, ldg, P, dx_lo, lddx_lo, dx_hi, lddx_hi, dw, accumulate, part, st);
        default: return launch_bwd<3>(x_lo, ld_lo, x_hi, ld_hi, w, g_agg, ldg, P, dx_lo, lddx_lo, dx_hi, lddx_hi, dw, accumulate, part, st);
    }
}

int pu_att16_bwd(const float *feature_set, int ldx, const float *w, const float *g_agg, int ldg, long long P, float *dx,
                 int lddx, float *dw, int accumulate, void *workspace, size_t workspace_bytes, pu_stream_t stream) {
    if (!feature_set || !dx || ldx < D || lddx < D) return PU_ERR_INVALID_ARG;
    return pu_att16_bwd_split(feature_set, ldx, feature_set + D / 2, ldx, w, g_agg, ldg, P, dx, lddx, dx + D / 2, lddx, dw,
                              accumulate, workspace, workspace_bytes, stream);
}

}  // extern "C"
