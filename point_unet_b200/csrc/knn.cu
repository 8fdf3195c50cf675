// knn.cu -- exact batched K-nearest-neighbour search for 3-D point clouds on sm_100a.
//
// Replaces the reference's host path  DataProcessing.knn_search (PointSegment/helper_tool.py:84-94)
//   -> nearest_neighbors.knn_batch (utils/nearest_neighbors/knn.pyx:71-109)
//   -> cpp_knn_batch_omp (utils/nearest_neighbors/knn_.cxx:104-135; nanoflann kd-tree per cloud).
// Not a port of the kd-tree: a GPU-native bucketed search.
//
//   build   (per call, per cloud)  Morton-order the support points (30-bit code, in-tree LSD radix sort), cut the
//           sorted array into buckets of 32 consecutive points with tight AABBs, and group 32 buckets
//           into a super-bucket AABB.  The structure adapts to density by construction (every bucket
//           holds 32 points whether it lies in the dense organ blob or the sparse background).
//   search  one warp owns 32 Morton-consecutive queries (one per lane).  Each lane keeps its K best
//           (distance, index) pairs as sorted 64-bit keys IN REGISTERS.  The warp first scans the
//           buckets next to its own position (seed), then walks the super-buckets whose AABB lies
//           within the warp's current search radius, tests their buckets per lane against the lane's
//           own K-th distance, stages each surviving bucket (32 candidates, 512 B) in SHARED MEMORY
//           and lets all 32 lanes sweep it with broadcast reads.
//
// Exactness.  A bucket is skipped only when a conservative lower bound of the fp32 distance exceeds the
// current K-th distance.  The bound uses the same rounded operations as the distance itself, and fp32
// rounding is monotone, so bound <= computed distance for every point in the box: no neighbour (and no
// boundary tie) can be lost.  Distances are computed exactly like nanoflann's L2_Adaptor::evalMetric for
// dim 3 (nanoflann.hpp:343-346): d = q - p, ((dx*dx)+(dy*dy))+(dz*dz), each operation rounded to
// nearest, NO fused multiply-add (__fmul_rn/__fadd_rn/__fsub_rn are never contracted).
// Tie rule: ascending (distance, index) -- a total order, hence the result is independent of the visiting
// order and of scheduling (deterministic).
#include <cstdlib>
#include <float.h>

#include "common.cuh"

namespace pu {
namespace knn {

constexpr int BS = 32;              // points per bucket  (= one warp-wide candidate tile)
constexpr int SBS = 32;             // buckets per super-bucket
constexpr int SEARCH_WARPS = 4;     // warps per CTA in the search kernel
constexpr unsigned long long KEY_INIT = 0x7F800000FFFFFFFFull;  // (+inf, id 0xFFFFFFFF)
// Morton sort: stable LSD radix sort of (30-bit key, point index) pairs, one cloud per blockIdx.y, 3 passes of 10 bits
constexpr int SORT_BITS = 10, SORT_BINS = 1 << SORT_BITS, SORT_PASSES = 3;
constexpr int SORT_THREADS = 256, SORT_IPT = 16, SORT_TILE = SORT_THREADS * SORT_IPT;  // 4096 pairs per CTA
constexpr int SORT_FUSED_TILES = 64;  // up to this many tiles per cloud the scatter CTAs scan the histogram themselves

struct Layout {
    size_t stats, bbox, keys_a, keys_b, vals_a, vals_b, sp, bk_lo, bk_hi, sb_lo, sb_hi;
    size_t qkeys_a, qkeys_b, qvals_a, qvals_b, sq, sort_hist, hist_one, cb_lo, cb_hi, total;
    int NB, NSB, NCB;
};

static Layout make_layout(int B, int N1, int N2) {
    Layout L;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    const size_t n1 = (size_t)B * N1, n2 = (size_t)B * N2;
    L.NB = ceil_div(N1, BS);
    L.NSB = ceil_div(L.NB, SBS);
    L.NCB = ceil_div(L.NSB, SBS);   // third level (32 super-buckets = 32 768 points): used by the warp-per-query search
    L.stats = take(8 * sizeof(unsigned long long));
    L.bbox = take((size_t)B * 8 * sizeof(unsigned));
    L.keys_a = take(n1 * 4);
    L.keys_b = take(n1 * 4);
    L.vals_a = take(n1 * 4);
    L.vals_b = take(n1 * 4);
    L.sp = take(n1 * 16);
    L.bk_lo = take((size_t)B * L.NB * 16);
    L.bk_hi = take((size_t)B * L.NB * 16);
    L.sb_lo = take((size_t)B * L.NSB * 16);
    L.sb_hi = take((size_t)B * L.NSB * 16);
    L.qkeys_a = take(n2 * 4);
    L.qkeys_b = take(n2 * 4);
    L.qvals_a = take(n2 * 4);
    L.qvals_b = take(n2 * 4);
    L.sq = take(n2 * 16);
    L.hist_one = align_up((size_t)B * SORT_BINS * ceil_div(N1 > N2 ? N1 : N2, SORT_TILE) * sizeof(unsigned), 256);
    L.sort_hist = take(L.hist_one * 2 * SORT_PASSES);   // [support | query][pass] digit histograms, zeroed by one memset
    L.cb_lo = take((size_t)B * L.NCB * 16);
    L.cb_hi = take((size_t)B * L.NCB * 16);
    L.total = off;
    return L;
}

// ---------------------------------------------------------------------------------------------
// order-preserving float <-> uint encoding (for atomicMin/Max bounding boxes)
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void bbox_init_kernel(unsigned *bbox, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 8) bbox[i] = ((i & 7) < 3) ? 0xFFFFFFFFu : 0u;  // [0..2] = min (start high), [4..6] = max
}

// grid (x, B): per-cloud bounding box of the support points
__global__ void __launch_bounds__(256) bbox_kernel(const float *__restrict__ pts, int N, unsigned *__restrict__ bbox) {
    const int b = blockIdx.y;
    const float *p = pts + (size_t)b * N * 3;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = p[(size_t)i * 3 + c];
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    }
    // one atomic per CTA and bound (every warp hitting the same six words serialises in L2: 74 us at 4 x 180 k points)
    __shared__ float s_lo[8][3], s_hi[8][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { s_lo[warp][c] = lo[c]; s_hi[warp][c] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        float l = s_lo[0][c], h = s_hi[0][c];
#pragma unroll
        for (int w = 1; w < 8; ++w) { l = fminf(l, s_lo[w][c]); h = fmaxf(h, s_hi[w][c]); }
        atomicMin(&bbox[b * 8 + c], f2ord(l));
        atomicMax(&bbox[b * 8 + 4 + c], f2ord(h));
    }
}

__device__ __forceinline__ unsigned spread10(unsigned v) {  // 10 bits -> every third bit
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// key = morton30(point quantised in the SUPPORT cloud's bounding box); val = row inside the cloud
__global__ void __launch_bounds__(256) morton_kernel(const float *__restrict__ pts, int N, int B,
                                                     const unsigned *__restrict__ bbox,
                                                     unsigned *__restrict__ keys,
                                                     unsigned *__restrict__ vals, unsigned *__restrict__ hist0, int ntiles) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)B * N) return;
    const int b = (int)(g / N);
    unsigned q[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float lo = ord2f(bbox[b * 8 + c]), hi = ord2f(bbox[b * 8 + 4 + c]);
        const float ext = hi - lo;
        const float scale = ext > 0.f ? 1024.f / ext : 0.f;
        float t = (pts[g * 3 + c] - lo) * scale;
        t = fminf(fmaxf(t, 0.f), 1023.f);  // also maps NaN to 0
        q[c] = (unsigned)t;
    }
    const unsigned m = (spread10(q[0]) << 2) | (spread10(q[1]) << 1) | spread10(q[2]);
    const unsigned i = (unsigned)(g - (size_t)b * N);
    keys[g] = m;
    vals[g] = i;
    // digit histogram of the first sort pass, per 4096-pair tile (integer atomics: the counts do not depend on their order)
    atomicAdd(&hist0[((size_t)b * ntiles + i / SORT_TILE) * SORT_BINS + (m & (SORT_BINS - 1))], 1u);
}

// ---------------------------------------------------------------------------------------------
// In-tree radix sort (replaces nanoflann's divideTree, nanoflann.hpp:916-964, as the spatial ordering step).  Per pass:
//   histograms           [cloud][tile][digit] counts of every pass: pass 0 is accumulated by morton_kernel, pass p+1 by the
//                        scatter kernel of pass p at the pairs' new positions (integer atomics, order-free)
//   sort_scan_kernel     (only for > SORT_FUSED_TILES tiles) exclusive scan over (digit, tile) = first output slot of every
//                        (digit, tile) group; otherwise every scatter CTA derives its own slots from the raw counts
//   sort_scatter_kernel  CTA (tile, cloud): STABLE ranks.  Warp w owns 512 consecutive pairs and walks them 32 at a time;
//                        inside a round the rank among equal digits comes from __match_any_sync (lower lanes first), across
//                        rounds and warps from per-warp digit counters that start at the scanned offset -- so equal keys keep
//                        their input order, every pass is a permutation that depends only on the data (deterministic).
// many tiles (> SORT_FUSED_TILES): exclusive scan over (digit, tile) as its own launch.  One CTA per cloud, thread d owns
// digit d's row of tile counts.
__global__ void __launch_bounds__(SORT_BINS) sort_scan_kernel(unsigned *__restrict__ hist, int ntiles) {
    __shared__ unsigned s_w[SORT_BINS / 32];
    unsigned *col = hist + (size_t)blockIdx.x * ntiles * SORT_BINS + threadIdx.x;   // [tile][digit]: coalesced over digits
    unsigned total = 0;
    for (int t = 0; t < ntiles; ++t) total += col[(size_t)t * SORT_BINS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned v = s_w[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += u;
        }
        s_w[lane] = iv - v;
    }
    __syncthreads();
    unsigned run = s_w[w] + inc - total;
    for (int t = 0; t < ntiles; ++t) {
        const unsigned c = col[(size_t)t * SORT_BINS];
        col[(size_t)t * SORT_BINS] = run;
        run += c;
    }
}

// SCANNED = false: `hist` holds raw counts [tile][digit] and every CTA derives its own first slots (few tiles: the whole
// matrix is a few hundred KB of L2 reads per CTA, cheaper than a scan launch); SCANNED = true: sort_scan_kernel ran before.
// `hist_next` (optional) receives the digit histogram of the NEXT pass at the pairs' new positions.
template <bool SCANNED>
__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const unsigned *__restrict__ keys_in,
                                                                    const unsigned *__restrict__ vals_in, int N, int ntiles,
                                                                    int shift, const unsigned *__restrict__ hist,
                                                                    unsigned *__restrict__ keys_out, unsigned *__restrict__ vals_out,
                                                                    unsigned *__restrict__ hist_next) {
    constexpr int WARPS = SORT_THREADS / 32, ROUNDS = SORT_TILE / SORT_THREADS;  // 8 warps x 16 rounds of 32 pairs
    constexpr int DPT = SORT_BINS / SORT_THREADS;                                // digits per thread (consecutive)
    __shared__ unsigned s_c[WARPS][SORT_BINS];
    __shared__ unsigned s_w[WARPS];
    const int b = blockIdx.y, tile = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < WARPS * SORT_BINS; i += SORT_THREADS) (&s_c[0][0])[i] = 0;
    __syncthreads();
    const size_t cb = (size_t)b * N;
    const int base = tile * SORT_TILE + w * (ROUNDS * 32);
    unsigned key[ROUNDS], val[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int i = base + r * 32 + lane;
        key[r] = 0; val[r] = 0;
        if (i < N) {
            key[r] = keys_in[cb + i];
            val[r] = vals_in[cb + i];
            atomicAdd(&s_c[w][(key[r] >> shift) & (SORT_BINS - 1)], 1u);   // counts only: order-free
        }
    }
    // first output slot of (digit, this tile) for the thread's DPT consecutive digits
    const unsigned *h = hist + (size_t)b * ntiles * SORT_BINS + threadIdx.x * DPT;   // [tile][digit], 4 digits = one 16-byte load
    static_assert(DPT == 4, "one uint4 per thread and tile");
    unsigned first[DPT];
    if (SCANNED) {
        const uint4 v = *reinterpret_cast<const uint4 *>(h + (size_t)tile * SORT_BINS);
        first[0] = v.x; first[1] = v.y; first[2] = v.z; first[3] = v.w;
        __syncthreads();
    } else {
        unsigned tot[DPT] = {0, 0, 0, 0}, mine = 0;
#pragma unroll
        for (int j = 0; j < DPT; ++j) first[j] = 0;
        for (int t = 0; t < ntiles; ++t) {
            const uint4 v = *reinterpret_cast<const uint4 *>(h + (size_t)t * SORT_BINS);
            const unsigned c[DPT] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                tot[j] += c[j];
                if (t < tile) first[j] += c[j];
            }
        }
#pragma unroll
        for (int j = 0; j < DPT; ++j) mine += tot[j];
        unsigned inc = mine;   // block-wide exclusive scan of the per-thread totals (digits ascend with the thread id)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        unsigned run = inc - mine;
#pragma unroll
        for (int i = 0; i < WARPS; ++i) run += i < w ? s_w[i] : 0u;
#pragma unroll
        for (int j = 0; j < DPT; ++j) { first[j] += run; run += tot[j]; }
    }
#pragma unroll
    for (int j = 0; j < DPT; ++j) {   // counters -> first slot of (digit, warp)
        const int d = threadIdx.x * DPT + j;
        unsigned run = first[j];
#pragma unroll
        for (int ww = 0; ww < WARPS; ++ww) {
            const unsigned c = s_c[ww][d];
            s_c[ww][d] = run;
            run += c;
        }
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1u;
    const int shift_next = shift + SORT_BITS;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int i = base + r * 32 + lane;
        const bool valid = i < N;
        const unsigned d = valid ? ((key[r] >> shift) & (SORT_BINS - 1)) : (unsigned)SORT_BINS;  // padding lanes form their own group
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        unsigned slot = 0;
        if (valid && lane == leader) {
            slot = s_c[w][d];
            s_c[w][d] = slot + __popc(peers);
        }
        slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(peers & lt);
        if (valid) {
            keys_out[cb + slot] = key[r];
            vals_out[cb + slot] = val[r];
            if (hist_next)
                atomicAdd(&hist_next[((size_t)b * ntiles + slot / SORT_TILE) * SORT_BINS + ((key[r] >> shift_next) & (SORT_BINS - 1))], 1u);
        }
        __syncwarp();
    }
}

// sorted[g] = (x, y, z, bits(local index))
__global__ void __launch_bounds__(256) gather_sorted_kernel(const float *__restrict__ pts, int N, int B,
                                                            const unsigned *__restrict__ vals_sorted,
                                                            float4 *__restrict__ out) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)B * N) return;
    const unsigned src = vals_sorted[g];  // row inside the cloud
    const int b = (int)(g / N);
    const float *p = pts + ((size_t)b * N + src) * 3;
    out[g] = make_float4(p[0], p[1], p[2], __int_as_float((int)src));
}

// one warp per group of 32 consecutive items: AABB of the group
// level 0: items are points of sp (valid count from N); level 1: items are bucket boxes
__global__ void __launch_bounds__(128) bucket_box_kernel(const float4 *__restrict__ sp, int N, int NB, int B,
                                                         float4 *__restrict__ lo_out, float4 *__restrict__ hi_out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= B * NB) return;
    const int b = w / NB, t = w - b * NB;
    const int i = t * BS + lane;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < N) {
        const float4 p = sp[(size_t)b * N + i];
        lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    if (lane == 0) {
        lo_out[w] = make_float4(lo[0], lo[1], lo[2], 0.f);
        hi_out[w] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
}

__global__ void __launch_bounds__(128) super_box_kernel(const float4 *__restrict__ bk_lo, const float4 *__restrict__ bk_hi,
                                                        int NB, int NSB, int B, float4 *__restrict__ lo_out,
                                                        float4 *__restrict__ hi_out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= B * NSB) return;
    const int b = w / NSB, s = w - b * NSB;
    const int t = s * SBS + lane;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (t < NB) {
        const float4 l = bk_lo[(size_t)b * NB + t], h = bk_hi[(size_t)b * NB + t];
        lo[0] = l.x; lo[1] = l.y; lo[2] = l.z;
        hi[0] = h.x; hi[1] = h.y; hi[2] = h.z;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    if (lane == 0) {
        lo_out[w] = make_float4(lo[0], lo[1], lo[2], 0.f);
        hi_out[w] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 arithmetic that must round like the reference (no FMA contraction, ever)
__device__ __forceinline__ float dist2_rn(float qx, float qy, float qz, float px, float py, float pz) {
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
// conservative lower bound of dist2_rn(q, p) over all p in [lo, hi]
__device__ __forceinline__ float point_box_dist2(float qx, float qy, float qz, const float4 &lo, const float4 &hi) {
    const float dx = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.f);
    const float dz = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.f);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
// conservative lower bound over all q in [qlo, qhi], p in [lo, hi]
__device__ __forceinline__ float box_box_dist2(const float (&qlo)[3], const float (&qhi)[3], const float4 &lo,
                                               const float4 &hi) {
    const float dx = fmaxf(fmaxf(__fsub_rn(lo.x, qhi[0]), __fsub_rn(qlo[0], hi.x)), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rn(lo.y, qhi[1]), __fsub_rn(qlo[1], hi.y)), 0.f);
    const float dz = fmaxf(fmaxf(__fsub_rn(lo.z, qhi[2]), __fsub_rn(qlo[2], hi.z)), 0.f);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// sorted insertion of key x (precondition: x < best[K-1]) into the ascending register list
template <int K>
__device__ __forceinline__ void topk_insert(unsigned long long (&best)[K], unsigned long long x) {
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
        const unsigned long long prev = best[j - 1];
        best[j] = (x < prev) ? prev : ((x < best[j]) ? x : best[j]);
    }
    best[0] = (x < best[0]) ? x : best[0];
}

template <int K>
struct WarpSearch {
    unsigned long long best[K];
    float qx, qy, qz;
    bool valid;
    unsigned long long n_evals;
    unsigned n_buckets, n_tests;

    __device__ __forceinline__ float kth() const { return __uint_as_float((unsigned)(best[K - 1] >> 32)); }

    // all 32 lanes sweep the `cnt` candidates of one bucket staged in shared memory (broadcast reads).
    // Measured alternatives that were SLOWER on B200 (round 2, 180k uniform cloud, this kernel 0.55 ms): queueing accepted
    // keys per lane and draining the queues together (15 % fewer instructions, 23 % more time: the drains serialise), and
    // per-lane bucket sweeps (3.1x fewer distance evaluations, 467 instead of 1462 per query, but 37 % more time: a 16-byte
    // load per lane and candidate instead of one broadcast) -- ncu: 62 % of the instructions are the sorted insertion.
    __device__ __forceinline__ void sweep_bucket(const float4 *__restrict__ sp_cloud, int N1, int t, float4 *tile,
                                                 int lane) {
        const int base = t * BS;
        const int cnt = min(BS, N1 - base);
        __syncwarp();
        if (lane < cnt) tile[lane] = sp_cloud[base + lane];
        __syncwarp();
        if (valid) {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const float4 p = tile[j];  // broadcast read
                const float d = dist2_rn(qx, qy, qz, p.x, p.y, p.z);
                const unsigned long long key =
                    ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
                if (key < best[K - 1]) topk_insert<K>(best, key);
            }
        }
        n_evals += cnt;
        n_buckets += 1;
    }
};

// grid: (ceil(warps_per_cloud / SEARCH_WARPS), B); one warp = 32 Morton-consecutive queries
template <int K>
__global__ void __launch_bounds__(SEARCH_WARPS * 32)
    knn_search_kernel(const float4 *__restrict__ sp, const float4 *__restrict__ sq,
                      const float4 *__restrict__ bk_lo, const float4 *__restrict__ bk_hi,
                      const float4 *__restrict__ sb_lo, const float4 *__restrict__ sb_hi,
                      const unsigned *__restrict__ skeys,  // sorted support keys (NULL for self-query)
                      const unsigned *__restrict__ qkeys,  // sorted query keys   (NULL for self-query)
                      int N1, int N2, int NB, int NSB, int kout, int32_t *__restrict__ out_idx,
                      float *__restrict__ out_dist, unsigned long long *__restrict__ stats) {
    __shared__ float4 s_tile[SEARCH_WARPS][BS];
    __shared__ float4 s_blo[SEARCH_WARPS][SBS];
    __shared__ float4 s_bhi[SEARCH_WARPS][SBS];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * SEARCH_WARPS + wib;  // warp index within the cloud
    const int b = blockIdx.y;
    const int nwarps = (N2 + 31) >> 5;
    if (w >= nwarps) return;  // whole warp exits together; no block-wide barriers are used below

    const float4 *sp_cloud = sp + (size_t)b * N1;
    const float4 *bk_lo_c = bk_lo + (size_t)b * NB, *bk_hi_c = bk_hi + (size_t)b * NB;
    const float4 *sb_lo_c = sb_lo + (size_t)b * NSB, *sb_hi_c = sb_hi + (size_t)b * NSB;
    float4 *tile = s_tile[wib];

    WarpSearch<K> S;
#pragma unroll
    for (int j = 0; j < K; ++j) S.best[j] = KEY_INIT;
    S.n_evals = 0; S.n_buckets = 0; S.n_tests = 0;

    const int qi = w * 32 + lane;
    S.valid = qi < N2;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (S.valid) q = sq[(size_t)b * N2 + qi];
    S.qx = q.x; S.qy = q.y; S.qz = q.z;
    const int q_orig = __float_as_int(q.w);

    // AABB of the warp's queries
    float qlo[3], qhi[3];
    qlo[0] = S.valid ? q.x : FLT_MAX; qhi[0] = S.valid ? q.x : -FLT_MAX;
    qlo[1] = S.valid ? q.y : FLT_MAX; qhi[1] = S.valid ? q.y : -FLT_MAX;
    qlo[2] = S.valid ? q.z : FLT_MAX; qhi[2] = S.valid ? q.z : -FLT_MAX;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            qlo[c] = fminf(qlo[c], __shfl_xor_sync(0xffffffffu, qlo[c], o));
            qhi[c] = fmaxf(qhi[c], __shfl_xor_sync(0xffffffffu, qhi[c], o));
        }

    // ---- seed: the buckets around the warp's own position in the support's Morton order
    int home;
    if (skeys == nullptr) {
        home = w;  // self-query: bucket w holds exactly these 32 points
    } else {
        const int nvalid = min(32, N2 - w * 32);
        int pos = 0;
        if (lane == 0) {
            const unsigned key = qkeys[(size_t)b * N2 + w * 32 + (nvalid - 1) / 2];
            const unsigned *sk = skeys + (size_t)b * N1;
            int lo = 0, hi = N1;  // lower_bound
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (sk[mid] < key) lo = mid + 1; else hi = mid;
            }
            pos = lo;
        }
        pos = __shfl_sync(0xffffffffu, pos, 0);
        home = min(pos, N1 - 1) / BS;
    }
    home = min(home, NB - 1);
    const int seed_lo = max(home - 1, 0), seed_hi = min(home + 1, NB - 1);
    S.sweep_bucket(sp_cloud, N1, home, tile, lane);
    for (int t = seed_lo; t <= seed_hi; ++t)
        if (t != home) S.sweep_bucket(sp_cloud, N1, t, tile, lane);

    // warp search radius^2 = max over lanes of the current K-th distance (+inf while a lane is not full)
    auto warp_radius = [&]() -> float {
        const unsigned bits = S.valid ? (unsigned)(S.best[K - 1] >> 32) : 0u;  // non-negative floats order as uints
        return __uint_as_float(__reduce_max_sync(0xffffffffu, bits));
    };
    float R2 = warp_radius();

    // ---- walk super-buckets, nearest chunk first (outward from the home chunk)
    const int nchunks = (NSB + 31) >> 5;
    const int home_chunk = (home / SBS) >> 5;
    for (int step = 0; step < 2 * nchunks; ++step) {  // step 0: home; 2k-1: home+k; 2k: home-k
        const int off = (step + 1) >> 1;
        const int chunk = (step & 1) ? home_chunk + off : home_chunk - off;
        if (chunk < 0 || chunk >= nchunks) continue;
        const int s = chunk * 32 + lane;
        float sd2 = FLT_MAX;
        bool hit = false;
        if (s < NSB) {
            sd2 = box_box_dist2(qlo, qhi, sb_lo_c[s], sb_hi_c[s]);
            hit = sd2 <= R2;
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        S.n_tests += 1;
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const float sd2b = __shfl_sync(0xffffffffu, sd2, bit);
            if (sd2b > R2) continue;  // radius shrank since the ballot
            const int sbi = chunk * 32 + bit;
            const int t0 = sbi * SBS;
            const int tn = min(SBS, NB - t0);
            // stage this super-bucket's bucket boxes, then every lane tests them against ITS OWN K-th distance
            __syncwarp();
            if (lane < tn) {
                s_blo[wib][lane] = bk_lo_c[t0 + lane];
                s_bhi[wib][lane] = bk_hi_c[t0 + lane];
            }
            __syncwarp();
            unsigned want = 0;
            if (S.valid) {
                const float kd = S.kth();
#pragma unroll 4
                for (int j = 0; j < tn; ++j) {
                    const float bd = point_box_dist2(S.qx, S.qy, S.qz, s_blo[wib][j], s_bhi[wib][j]);
                    want |= (bd <= kd ? 1u : 0u) << j;
                }
            }
            unsigned need = __reduce_or_sync(0xffffffffu, want);
            S.n_tests += tn;
            while (need) {
                const int j = __ffs(need) - 1;
                need &= need - 1;
                const int t = t0 + j;
                if (t >= seed_lo && t <= seed_hi) continue;  // already swept
                // re-test against the lanes' current K-th distances (they shrink as buckets are swept)
                const bool still = S.valid && point_box_dist2(S.qx, S.qy, S.qz, s_blo[wib][j], s_bhi[wib][j]) <= S.kth();
                if (!__any_sync(0xffffffffu, still)) continue;
                S.sweep_bucket(sp_cloud, N1, t, tile, lane);
            }
            R2 = warp_radius();
        }
    }

    // ---- write back in the ORIGINAL query order
    if (S.valid) {
        int32_t *o = out_idx + ((size_t)b * N2 + q_orig) * kout;
        if ((kout & 3) == 0) {
#pragma unroll
            for (int j = 0; j < K; j += 4) {
                if (j < kout) {
                    int4 v;
                    v.x = S.best[j] == KEY_INIT ? 0 : (int)(unsigned)S.best[j];
                    v.y = S.best[(j + 1) % K] == KEY_INIT ? 0 : (int)(unsigned)S.best[(j + 1) % K];
                    v.z = S.best[(j + 2) % K] == KEY_INIT ? 0 : (int)(unsigned)S.best[(j + 2) % K];
                    v.w = S.best[(j + 3) % K] == KEY_INIT ? 0 : (int)(unsigned)S.best[(j + 3) % K];
                    *reinterpret_cast<int4 *>(o + j) = v;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j)
                if (j < kout) o[j] = S.best[j] == KEY_INIT ? 0 : (int)(unsigned)S.best[j];
        }
        if (out_dist) {
            float *od = out_dist + ((size_t)b * N2 + q_orig) * kout;
#pragma unroll
            for (int j = 0; j < K; ++j)
                if (j < kout) od[j] = S.best[j] == KEY_INIT ? FLT_MAX : __uint_as_float((unsigned)(S.best[j] >> 32));
        }
    }
    if (stats) {
        unsigned long long ev = S.valid ? S.n_evals : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ev += __shfl_xor_sync(0xffffffffu, ev, o);
        if (lane == 0) {
            atomicAdd(&stats[0], ev);
            atomicAdd(&stats[1], (unsigned long long)S.n_buckets);
            atomicAdd(&stats[2], (unsigned long long)S.n_tests);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Warp-per-query search (self-queries, K <= 32).  The per-lane kernel above makes 32 queries share every candidate
// bucket, so each lane evaluates the union of 32 neighbourhoods (~1 460 distances per query at 180 k points) and the warp
// runs the 16-deep sorted insertion whenever ANY lane accepts a candidate -- i.e. almost always.  Here a warp owns ONE
// query: lane l evaluates candidate l of a bucket (one coalesced 512-byte load), the K best keys live as an ascending list
// ACROSS the lanes, a candidate is inserted with one ballot + one shuffle-up, many candidates at once by a warp bitonic sort
// + bitonic merge, and boxes are pruned against the query's OWN K-th distance on three levels (32 768 / 1 024 / 32 points).
// Same distance arithmetic, same conservative bounds, same (distance, index) total order => bit-identical results.
constexpr int QW_WARPS = 8;      // warps per CTA
constexpr int QW_QPW = 8;        // queries per warp: a CTA covers 64 Morton-consecutive queries (shared L1 working set)
constexpr int QW_MERGE_MIN = 12;  // more accepted candidates than this in one bucket: sort + merge instead of insertions

__device__ __forceinline__ unsigned long long qw_sort32(unsigned long long x, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, x, j);
            const bool keep_min = (((lane & k) == 0) == ((lane & j) == 0));
            x = ((x < o) == keep_min) ? x : o;
        }
    return x;
}
// `list` and `y` ascending by lane -> the 32 smallest keys of their union, ascending by lane
__device__ __forceinline__ unsigned long long qw_merge32(unsigned long long list, unsigned long long y, int lane) {
    const unsigned long long yr = __shfl_sync(0xffffffffu, y, 31 - lane);
    unsigned long long z = list < yr ? list : yr;  // bitonic
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, z, j);
        const bool keep_min = (lane & j) == 0;
        z = ((z < o) == keep_min) ? z : o;
    }
    return z;
}

__global__ void __launch_bounds__(QW_WARPS * 32)
    knn_query_warp_kernel(const float4 *__restrict__ sp, const float4 *__restrict__ bk_lo, const float4 *__restrict__ bk_hi,
                          const float4 *__restrict__ sb_lo, const float4 *__restrict__ sb_hi,
                          const float4 *__restrict__ cb_lo, const float4 *__restrict__ cb_hi, int N, int NB, int NSB,
                          int NCB, int kout, int32_t *__restrict__ out_idx, float *__restrict__ out_dist,
                          unsigned long long *__restrict__ stats) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, b = blockIdx.y;
    const float4 *sp_cloud = sp + (size_t)b * N;
    const float4 *bk_lo_c = bk_lo + (size_t)b * NB, *bk_hi_c = bk_hi + (size_t)b * NB;
    const float4 *sb_lo_c = sb_lo + (size_t)b * NSB, *sb_hi_c = sb_hi + (size_t)b * NSB;
    const float4 *cb_lo_c = cb_lo + (size_t)b * NCB, *cb_hi_c = cb_hi + (size_t)b * NCB;
    unsigned long long n_evals = 0;
    unsigned n_buckets = 0, n_tests = 0;

    for (int i = 0; i < QW_QPW; ++i) {
        const int qi = (blockIdx.x * QW_QPW + i) * QW_WARPS + wib;  // warp-uniform
        if (qi >= N) break;
        const float4 q = sp_cloud[qi];
        unsigned long long list = KEY_INIT, thr = KEY_INIT;  // ascending across the lanes; thr = the kout-th key
        unsigned q_evals = 0;
        float kd = __uint_as_float(0x7F800000u);

        // candidate keys of bucket t, one per lane (KEY_INIT beyond the end of the cloud)
        auto load_keys = [&](int t) -> unsigned long long {
            const int base = t * BS;
            const int cnt = min(BS, N - base);
            unsigned long long key = KEY_INIT;
            if (lane < cnt) {
                const float4 p = sp_cloud[base + lane];
                const float d = dist2_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
            }
            q_evals += cnt;
            return key;
        };
        auto set_thr = [&]() {
            thr = __shfl_sync(0xffffffffu, list, kout - 1);
            kd = __uint_as_float((unsigned)(thr >> 32));
        };
        // fold one bucket's keys into the list
        auto absorb = [&](unsigned long long key) {
            unsigned m = __ballot_sync(0xffffffffu, key < thr);
            if (m == 0) return;
            if (__popc(m) > QW_MERGE_MIN) {
                list = qw_merge32(list, qw_sort32(key, lane), lane);
            } else {
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned long long x = __shfl_sync(0xffffffffu, key, bit);
                    // position = number of smaller keys; everything from there on moves up one lane
                    const int pos = __popc(__ballot_sync(0xffffffffu, list < x));
                    const unsigned long long up = __shfl_up_sync(0xffffffffu, list, 1);
                    list = lane < pos ? list : (lane == pos ? x : up);
                }
            }
            set_thr();
        };

        // seed: the query's own bucket and its two Morton neighbours (three loads in flight)
        const int home = min(qi / BS, NB - 1);
        const int seed_lo = max(home - 1, 0), seed_hi = min(home + 1, NB - 1);
        {
            const unsigned long long k0 = load_keys(home);
            const unsigned long long k1 = seed_lo != home ? load_keys(seed_lo) : KEY_INIT;
            const unsigned long long k2 = seed_hi != home ? load_keys(seed_hi) : KEY_INIT;
            list = qw_sort32(k0, lane);
            set_thr();
            absorb(k1);
            absorb(k2);
        }

        // Boxes are visited NEAREST FIRST on the two lower levels (one REDUX over the lanes' bound bits picks the next box;
        // non-negative floats order like their bit patterns), so the radius shrinks early and the walk of a level stops at
        // the first remaining box that lies beyond it.
        for (int c0 = 0; c0 < NCB; c0 += 32) {
            float cd = FLT_MAX;
            const bool c_ok = c0 + lane < NCB;   // (kd is +inf while the list is not full: validity is its own predicate)
            if (c_ok) cd = point_box_dist2(q.x, q.y, q.z, cb_lo_c[c0 + lane], cb_hi_c[c0 + lane]);
            unsigned mc = __ballot_sync(0xffffffffu, c_ok && cd <= kd);
            n_tests += 1;
            while (mc) {
                const int bc = __ffs(mc) - 1;
                mc &= mc - 1;
                if (__shfl_sync(0xffffffffu, cd, bc) > kd) continue;  // the radius shrank since the ballot
                const int s0 = (c0 + bc) * SBS;
                unsigned sbits = 0xFFFFFFFFu;   // bound of this lane's super-bucket; all ones = nothing (left) to visit
                if (s0 + lane < NSB)
                    sbits = __float_as_uint(point_box_dist2(q.x, q.y, q.z, sb_lo_c[s0 + lane], sb_hi_c[s0 + lane]));
                n_tests += 1;
                for (;;) {
                    const unsigned smin = __reduce_min_sync(0xffffffffu, sbits);
                    if (smin == 0xFFFFFFFFu || __uint_as_float(smin) > kd) break;
                    const int bs = __ffs(__ballot_sync(0xffffffffu, sbits == smin)) - 1;
                    if (lane == bs) sbits = 0xFFFFFFFFu;
                    const int t0 = (s0 + bs) * SBS, t = t0 + lane;
                    unsigned bbits = 0xFFFFFFFFu;
                    if (t < NB && (t < seed_lo || t > seed_hi))   // the seed buckets were swept already
                        bbits = __float_as_uint(point_box_dist2(q.x, q.y, q.z, bk_lo_c[t], bk_hi_c[t]));
                    n_tests += 1;
                    for (;;) {
                        const unsigned bmin = __reduce_min_sync(0xffffffffu, bbits);
                        if (bmin == 0xFFFFFFFFu || __uint_as_float(bmin) > kd) break;
                        const int bb = __ffs(__ballot_sync(0xffffffffu, bbits == bmin)) - 1;
                        if (lane == bb) bbits = 0xFFFFFFFFu;
                        absorb(load_keys(t0 + bb));   // (fetching the two nearest buckets together bought nothing: issue-bound)
                    }
                }
            }
        }

        n_evals += q_evals;
        n_buckets += (q_evals + BS - 1) / BS;
        if (lane < kout) {
            const size_t o = ((size_t)b * N + (size_t)__float_as_int(q.w)) * kout + lane;
            out_idx[o] = list == KEY_INIT ? 0 : (int)(unsigned)list;
            if (out_dist) out_dist[o] = list == KEY_INIT ? FLT_MAX : __uint_as_float((unsigned)(list >> 32));
        }
    }
    if (stats && lane == 0) {
        atomicAdd(&stats[0], n_evals);
        atomicAdd(&stats[1], (unsigned long long)n_buckets);
        atomicAdd(&stats[2], (unsigned long long)n_tests);
    }
}

// ---------------------------------------------------------------------------------------------
// Up-sampling indices from the neighbour lists (tf_map: interp_idx = knn(sub_points, points, 1) with sub_points = the
// first n_sub points, runPancreas.py:135-137).  The nearest sub-cloud point of point n is the FIRST entry of n's own
// K-neighbour row whose index is below n_sub: the row holds the K smallest (distance, index) keys among ALL points in
// ascending order, so every sub-cloud point outside the row has a larger key than every point inside it.  Only rows without
// such an entry (about 0.75^16 = 1 % of them at a sub-sampling ratio of 4) need a search; they are collected in a list and
// searched, one warp per row, on the structure the self-query has just built -- candidates filtered by index < n_sub, same
// distance arithmetic, same bounds, same (distance, index) order, hence bit-identical to a separate K = 1 search over the
// sub-cloud.  That search used to cost a Morton sort of the sub-cloud, one of all query points and a full K = 1 sweep per level.
// smallest ORIGINAL point index inside every box of the three levels (one warp per box: 32 points, 32 buckets, 32
// super-buckets).  The filtered search below skips a box that holds no point of the prefix at all; without this, a cloud
// whose first n_sub points are a spatial REGION rather than a random subset (a volume listed organ first, say) sends most
// rows into the filtered search AND makes every such row wade through all the non-prefix points that lie nearer (measured:
// 94 ms instead of 1.6 ms per level).
__global__ void __launch_bounds__(128) box_min_idx_kernel(const float4 *__restrict__ sp, const int *__restrict__ child, int n_items,
                                                          int n_boxes, int B, int *__restrict__ out) {
    const long long gw = ((long long)blockIdx.x * 128 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)B * n_boxes) return;
    const int b = (int)(gw / n_boxes), t = (int)(gw % n_boxes);
    const int i = t * 32 + lane;
    unsigned v = 0xFFFFFFFFu;
    if (i < n_items) v = sp ? (unsigned)__float_as_int(sp[(size_t)b * n_items + i].w) : (unsigned)child[(size_t)b * n_items + i];
    v = __reduce_min_sync(0xffffffffu, v);
    if (lane == 0) out[(size_t)b * n_boxes + t] = (int)v;
}

__global__ void __launch_bounds__(256) interp_from_neigh_kernel(const int32_t *__restrict__ neigh, int N, int K, int kvalid,
                                                                int n_sub, int32_t *__restrict__ interp,
                                                                unsigned *__restrict__ cnt, unsigned *__restrict__ list) {
    const int b = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    const int32_t *row = neigh + ((size_t)b * N + n) * K;
    int found = -1;
    for (int k = 0; k < kvalid; ++k) {
        const int j = row[k];
        if (j < n_sub) { found = j; break; }
    }
    if (found >= 0) interp[(size_t)b * N + n] = found;
    else list[(size_t)b * N + atomicAdd(&cnt[b], 1u)] = (unsigned)n;   // the ORDER of the list is immaterial
}

constexpr int NP_WARPS = 8;
__global__ void __launch_bounds__(NP_WARPS * 32)
    knn_nearest_prefix_kernel(const float *__restrict__ cloud, const float4 *__restrict__ sp, const float4 *__restrict__ bk_lo,
                              const float4 *__restrict__ bk_hi, const float4 *__restrict__ sb_lo,
                              const float4 *__restrict__ sb_hi, const float4 *__restrict__ cb_lo,
                              const float4 *__restrict__ cb_hi, const int *__restrict__ bk_min,
                              const int *__restrict__ sb_min, const int *__restrict__ cb_min, int N, int NB, int NSB, int NCB,
                              int n_sub, const unsigned *__restrict__ cnt, const unsigned *__restrict__ list,
                              int32_t *__restrict__ interp) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, b = blockIdx.y;
    const float4 *sp_cloud = sp + (size_t)b * N;
    const float4 *bk_lo_c = bk_lo + (size_t)b * NB, *bk_hi_c = bk_hi + (size_t)b * NB;
    const float4 *sb_lo_c = sb_lo + (size_t)b * NSB, *sb_hi_c = sb_hi + (size_t)b * NSB;
    const float4 *cb_lo_c = cb_lo + (size_t)b * NCB, *cb_hi_c = cb_hi + (size_t)b * NCB;
    const int *bk_min_c = bk_min + (size_t)b * NB, *sb_min_c = sb_min + (size_t)b * NSB, *cb_min_c = cb_min + (size_t)b * NCB;
    const unsigned total = cnt[b];
    for (unsigned u = blockIdx.x * NP_WARPS + wib; u < total; u += gridDim.x * NP_WARPS) {   // warp-uniform
        const unsigned n = list[(size_t)b * N + u];
        const float *qp = cloud + ((size_t)b * N + n) * 3;
        const float qx = qp[0], qy = qp[1], qz = qp[2];
        unsigned long long best = KEY_INIT;
        float kd = __uint_as_float(0x7F800000u);
        // nearest first on all three levels; a box is skipped only when its lower bound EXCEEDS the best distance so far
        for (int c0 = 0; c0 < NCB; c0 += 32) {
            unsigned cbits = 0xFFFFFFFFu;
            if (c0 + lane < NCB && cb_min_c[c0 + lane] < n_sub)   // (boxes without a prefix point are never entered)
                cbits = __float_as_uint(point_box_dist2(qx, qy, qz, cb_lo_c[c0 + lane], cb_hi_c[c0 + lane]));
            for (;;) {
                const unsigned cmin = __reduce_min_sync(0xffffffffu, cbits);
                if (cmin == 0xFFFFFFFFu || __uint_as_float(cmin) > kd) break;
                const int bc = __ffs(__ballot_sync(0xffffffffu, cbits == cmin)) - 1;
                if (lane == bc) cbits = 0xFFFFFFFFu;
                const int s0 = (c0 + bc) * SBS;
                unsigned sbits = 0xFFFFFFFFu;
                if (s0 + lane < NSB && sb_min_c[s0 + lane] < n_sub)
                    sbits = __float_as_uint(point_box_dist2(qx, qy, qz, sb_lo_c[s0 + lane], sb_hi_c[s0 + lane]));
                for (;;) {
                    const unsigned smin = __reduce_min_sync(0xffffffffu, sbits);
                    if (smin == 0xFFFFFFFFu || __uint_as_float(smin) > kd) break;
                    const int bs = __ffs(__ballot_sync(0xffffffffu, sbits == smin)) - 1;
                    if (lane == bs) sbits = 0xFFFFFFFFu;
                    const int t0 = (s0 + bs) * SBS, t = t0 + lane;
                    unsigned bbits = 0xFFFFFFFFu;
                    if (t < NB && bk_min_c[t] < n_sub) bbits = __float_as_uint(point_box_dist2(qx, qy, qz, bk_lo_c[t], bk_hi_c[t]));
                    for (;;) {
                        const unsigned bmin = __reduce_min_sync(0xffffffffu, bbits);
                        if (bmin == 0xFFFFFFFFu || __uint_as_float(bmin) > kd) break;
                        const int bb = __ffs(__ballot_sync(0xffffffffu, bbits == bmin)) - 1;
                        if (lane == bb) bbits = 0xFFFFFFFFu;
                        const int base = (t0 + bb) * BS;
                        unsigned hi = 0xFFFFFFFFu, lo = 0xFFFFFFFFu;
                        if (base + lane < N) {
                            const float4 p = sp_cloud[base + lane];
                            const unsigned pi = (unsigned)__float_as_int(p.w);
                            if (pi < (unsigned)n_sub) {
                                hi = __float_as_uint(dist2_rn(qx, qy, qz, p.x, p.y, p.z));
                                lo = pi;
                            }
                        }
                        const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
                        if (mh == 0xFFFFFFFFu) continue;   // no sub-cloud point in this bucket
                        const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xFFFFFFFFu);
                        const unsigned long long cand = ((unsigned long long)mh << 32) | ml;
                        if (cand < best) { best = cand; kd = __uint_as_float(mh); }
                    }
                }
            }
        }
        if (lane == 0) interp[(size_t)b * N + n] = best == KEY_INIT ? 0 : (int)(unsigned)best;
    }
}

template <int K>
static int launch_search(const Layout &L, char *ws, bool self, const unsigned *skeys,
                         const unsigned *qkeys, const float4 *sq, int B, int N1, int N2, int kout,
                         int32_t *out_idx, float *out_dist, cudaStream_t st) {
    const int nwarps = ceil_div(N2, 32);
    dim3 grid(ceil_div(nwarps, SEARCH_WARPS), B);
    knn_search_kernel<K><<<grid, SEARCH_WARPS * 32, 0, st>>>(
        (const float4 *)(ws + L.sp), sq, (const float4 *)(ws + L.bk_lo), (const float4 *)(ws + L.bk_hi),
        (const float4 *)(ws + L.sb_lo), (const float4 *)(ws + L.sb_hi), self ? nullptr : skeys,
        self ? nullptr : qkeys, N1, N2, L.NB, L.NSB, kout, out_idx, out_dist,
        (unsigned long long *)(ws + L.stats));
    PU_LAUNCH_CHECK();
    return PU_OK;
}

// stable sort of the per-cloud (key, value) pairs by the low 30 bits of the key; `hist` = SORT_PASSES zeroed histogram
// buffers of `hist_one` bytes, the first already filled by morton_kernel.  The result lands in the "b" buffers (an odd
// number of passes), which *k_sorted / *v_sorted report.
static int sort_pairs(char *hist, size_t hist_one, unsigned *ka, unsigned *kb, unsigned *va, unsigned *vb, int B, int N,
                      cudaStream_t st, unsigned **k_sorted, unsigned **v_sorted) {
    const int ntiles = ceil_div(N, SORT_TILE);
    dim3 grid(ntiles, B);
    unsigned *kin = ka, *vin = va, *kout = kb, *vout = vb;
    for (int pass = 0; pass < SORT_PASSES; ++pass) {
        const int shift = pass * SORT_BITS;
        unsigned *h = (unsigned *)(hist + (size_t)pass * hist_one);
        unsigned *hn = pass + 1 < SORT_PASSES ? (unsigned *)(hist + (size_t)(pass + 1) * hist_one) : nullptr;
        if (ntiles > SORT_FUSED_TILES) {
            sort_scan_kernel<<<B, SORT_BINS, 0, st>>>(h, ntiles);
            PU_LAUNCH_CHECK();
            sort_scatter_kernel<true><<<grid, SORT_THREADS, 0, st>>>(kin, vin, N, ntiles, shift, h, kout, vout, hn);
        } else {
            sort_scatter_kernel<false><<<grid, SORT_THREADS, 0, st>>>(kin, vin, N, ntiles, shift, h, kout, vout, hn);
        }
        PU_LAUNCH_CHECK();
        unsigned *t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    *k_sorted = kin;
    *v_sorted = vin;
    return PU_OK;
}

static int knn_impl(const float *support, const float *query, int B, int N1, int N2, int K, int32_t *out_idx,
                    float *out_dist, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (!support || !query || !out_idx || B < 0 || N1 < 0 || N2 < 0 || K < 1 || K > PU_KNN_MAX_K)
        return PU_ERR_INVALID_ARG;
    if ((long long)B * N1 >= (1ll << 31) || (long long)B * N2 >= (1ll << 31) || B > (1 << 20)) return PU_ERR_UNSUPPORTED;
    if (B == 0 || N2 == 0) return PU_OK;
    if (N1 == 0) {  // the reference would abort (nanoflann.hpp:1331 throws); we return zeros like N1 < K
        PU_CUDA_TRY(cudaMemsetAsync(out_idx, 0, (size_t)B * N2 * K * sizeof(int32_t), st));
        return PU_OK;
    }
    const Layout L = make_layout(B, N1, N2);
    if (!workspace || workspace_bytes < L.total) return PU_ERR_WORKSPACE;
    char *ws = (char *)workspace;
    const bool self = (support == query) && (N1 == N2);
    const size_t n1 = (size_t)B * N1, n2 = (size_t)B * N2;
    unsigned *bbox = (unsigned *)(ws + L.bbox);

    PU_CUDA_TRY(cudaMemsetAsync(ws + L.stats, 0, 8 * sizeof(unsigned long long), st));
    PU_CUDA_TRY(cudaMemsetAsync(ws + L.sort_hist, 0, L.hist_one * (self ? 1 : 2) * SORT_PASSES, st));
    bbox_init_kernel<<<ceil_div(B * 8, 128), 128, 0, st>>>(bbox, B);
    PU_LAUNCH_CHECK();
    {
        dim3 grid(min(ceil_div(N1, 1024), kNumSMs), B);
        bbox_kernel<<<grid, 256, 0, st>>>(support, N1, bbox);
        PU_LAUNCH_CHECK();
    }
    morton_kernel<<<ceil_div(n1, 256), 256, 0, st>>>(support, N1, B, bbox, (unsigned *)(ws + L.keys_a),
                                                     (unsigned *)(ws + L.vals_a), (unsigned *)(ws + L.sort_hist),
                                                     ceil_div(N1, SORT_TILE));
    PU_LAUNCH_CHECK();
    unsigned *skeys = nullptr, *qkeys = nullptr;
    unsigned *svals = nullptr, *qvals = nullptr;
    int rc = sort_pairs(ws + L.sort_hist, L.hist_one, (unsigned *)(ws + L.keys_a), (unsigned *)(ws + L.keys_b),
                        (unsigned *)(ws + L.vals_a), (unsigned *)(ws + L.vals_b), B, N1, st, &skeys, &svals);
    if (rc != PU_OK) return rc;
    gather_sorted_kernel<<<ceil_div(n1, 256), 256, 0, st>>>(support, N1, B, svals, (float4 *)(ws + L.sp));
    PU_LAUNCH_CHECK();
    bucket_box_kernel<<<ceil_div((long long)B * L.NB * 32, 128), 128, 0, st>>>(
        (const float4 *)(ws + L.sp), N1, L.NB, B, (float4 *)(ws + L.bk_lo), (float4 *)(ws + L.bk_hi));
    PU_LAUNCH_CHECK();
    super_box_kernel<<<ceil_div((long long)B * L.NSB * 32, 128), 128, 0, st>>>(
        (const float4 *)(ws + L.bk_lo), (const float4 *)(ws + L.bk_hi), L.NB, L.NSB, B, (float4 *)(ws + L.sb_lo),
        (float4 *)(ws + L.sb_hi));
    PU_LAUNCH_CHECK();

    static const int warp_query = []() {
        const char *e = getenv("PU_KNN_WARP_QUERY");
        return e ? atoi(e) : 1;
    }();
    if (self && warp_query && K > 4) {  // one warp per query (see knn_query_warp_kernel)
        super_box_kernel<<<ceil_div((long long)B * L.NCB * 32, 128), 128, 0, st>>>(
            (const float4 *)(ws + L.sb_lo), (const float4 *)(ws + L.sb_hi), L.NSB, L.NCB, B, (float4 *)(ws + L.cb_lo),
            (float4 *)(ws + L.cb_hi));
        PU_LAUNCH_CHECK();
        dim3 grid(ceil_div(N1, QW_WARPS * QW_QPW), B);
        knn_query_warp_kernel<<<grid, QW_WARPS * 32, 0, st>>>(
            (const float4 *)(ws + L.sp), (const float4 *)(ws + L.bk_lo), (const float4 *)(ws + L.bk_hi),
            (const float4 *)(ws + L.sb_lo), (const float4 *)(ws + L.sb_hi), (const float4 *)(ws + L.cb_lo),
            (const float4 *)(ws + L.cb_hi), N1, L.NB, L.NSB, L.NCB, K, out_idx, out_dist,
            (unsigned long long *)(ws + L.stats));
        PU_LAUNCH_CHECK();
        return PU_OK;
    }

    const float4 *sq = (const float4 *)(ws + L.sp);
    if (!self) {
        char *qhist = ws + L.sort_hist + L.hist_one * SORT_PASSES;
        morton_kernel<<<ceil_div(n2, 256), 256, 0, st>>>(query, N2, B, bbox, (unsigned *)(ws + L.qkeys_a),
                                                         (unsigned *)(ws + L.qvals_a), (unsigned *)qhist,
                                                         ceil_div(N2, SORT_TILE));
        PU_LAUNCH_CHECK();
        rc = sort_pairs(qhist, L.hist_one, (unsigned *)(ws + L.qkeys_a), (unsigned *)(ws + L.qkeys_b),
                        (unsigned *)(ws + L.qvals_a), (unsigned *)(ws + L.qvals_b), B, N2, st, &qkeys, &qvals);
        if (rc != PU_OK) return rc;
        gather_sorted_kernel<<<ceil_div(n2, 256), 256, 0, st>>>(query, N2, B, qvals, (float4 *)(ws + L.sq));
        PU_LAUNCH_CHECK();
        sq = (const float4 *)(ws + L.sq);
    }

    int KT = 1;
    while (KT < K) KT <<= 1;
    switch (KT) {
        case 1: return launch_search<1>(L, ws, self, skeys, qkeys, sq, B, N1, N2, K, out_idx, out_dist, st);
        case 2: return launch_search<2>(L, ws, self, skeys, qkeys, sq, B, N1, N2, K, out_idx, out_dist, st);
        case 4: return launch_search<4>(L, ws, self, skeys, qkeys, sq, B, N1, N2, K, out_idx, out_dist, st);
        case 8: return launch_search<8>(L, ws, self, skeys, qkeys, sq, B, N1, N2, K, out_idx, out_dist, st);
        case 16: return launch_search<16>(L, ws, self, skeys, qkeys, sq, B, N1, N2, K, out_idx, out_dist, st);
        case 32: return launch_search<32>(L, ws, self, skeys, qkeys, sq, B, N1, N2, K, out_idx, out_dist, st);
    }
    return PU_ERR_UNSUPPORTED;
}

// neighbour lists of a cloud AND the up-sampling index of every point into the cloud's first n_sub points, from ONE
// search structure (one Morton sort per pyramid level instead of three)
static int knn_self_interp_impl(const float *cloud, int B, int N, int K, int n_sub, int32_t *out_neigh, int32_t *out_interp,
                                unsigned *out_unresolved, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (!cloud || !out_neigh || !out_interp || n_sub < 0 || n_sub > N) return PU_ERR_INVALID_ARG;
    int rc = knn_impl(cloud, cloud, B, N, N, K, out_neigh, nullptr, workspace, workspace_bytes, st);
    if (rc != PU_OK) return rc;
    if (B == 0 || N == 0) return PU_OK;
    if (n_sub == 0) {
        PU_CUDA_TRY(cudaMemsetAsync(out_interp, 0, (size_t)B * N * sizeof(int32_t), st));
        if (out_unresolved) PU_CUDA_TRY(cudaMemsetAsync(out_unresolved, 0, (size_t)B * sizeof(unsigned), st));
        return PU_OK;
    }
    const Layout L = make_layout(B, N, N);
    char *ws = (char *)workspace;
    // the query-side sort buffers are unused by a self-query: counters and row lists live there
    unsigned *cnt = (unsigned *)(ws + L.qkeys_a), *list = (unsigned *)(ws + L.qvals_a);
    if ((size_t)B * sizeof(unsigned) > (size_t)B * N * 4) return PU_ERR_WORKSPACE;
    PU_CUDA_TRY(cudaMemsetAsync(cnt, 0, (size_t)B * sizeof(unsigned), st));
    interp_from_neigh_kernel<<<dim3(ceil_div(N, 256), B), 256, 0, st>>>(out_neigh, N, K, K < N ? K : N, n_sub, out_interp, cnt,
                                                                        list);
    PU_LAUNCH_CHECK();
    // third box level (the warp-per-query self search builds it too; rebuilt here for the K <= 4 / per-lane path)
    super_box_kernel<<<ceil_div((long long)B * L.NCB * 32, 128), 128, 0, st>>>(
        (const float4 *)(ws + L.sb_lo), (const float4 *)(ws + L.sb_hi), L.NSB, L.NCB, B, (float4 *)(ws + L.cb_lo),
        (float4 *)(ws + L.cb_hi));
    PU_LAUNCH_CHECK();
    int *bk_min = (int *)(ws + L.qkeys_b), *sb_min = bk_min + (size_t)B * L.NB, *cb_min = sb_min + (size_t)B * L.NSB;
    box_min_idx_kernel<<<ceil_div((long long)B * L.NB * 32, 128), 128, 0, st>>>((const float4 *)(ws + L.sp), nullptr, N, L.NB, B,
                                                                               bk_min);
    PU_LAUNCH_CHECK();
    box_min_idx_kernel<<<ceil_div((long long)B * L.NSB * 32, 128), 128, 0, st>>>(nullptr, bk_min, L.NB, L.NSB, B, sb_min);
    PU_LAUNCH_CHECK();
    box_min_idx_kernel<<<ceil_div((long long)B * L.NCB * 32, 128), 128, 0, st>>>(nullptr, sb_min, L.NSB, L.NCB, B, cb_min);
    PU_LAUNCH_CHECK();
    int gx = ceil_div(N, NP_WARPS * 4);    // rows of the list per warp: few when the prefix is a random subset (~1 % of N)
    const int cap = kNumSMs * 8 / (B > 0 ? B : 1) + 1;
    if (gx > cap) gx = cap;
    knn_nearest_prefix_kernel<<<dim3(gx, B), NP_WARPS * 32, 0, st>>>(
        cloud, (const float4 *)(ws + L.sp), (const float4 *)(ws + L.bk_lo), (const float4 *)(ws + L.bk_hi),
        (const float4 *)(ws + L.sb_lo), (const float4 *)(ws + L.sb_hi), (const float4 *)(ws + L.cb_lo),
        (const float4 *)(ws + L.cb_hi), bk_min, sb_min, cb_min, N, L.NB, L.NSB, L.NCB, n_sub, cnt, list, out_interp);
    PU_LAUNCH_CHECK();
    if (out_unresolved)
        PU_CUDA_TRY(cudaMemcpyAsync(out_unresolved, cnt, (size_t)B * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    return PU_OK;
}

}  // namespace knn
}  // namespace pu

extern "C" {

size_t pu_knn_workspace_bytes(int B, int N1, int N2, int K) {
    (void)K;
    if (B <= 0 || N1 < 0 || N2 < 0) return 0;
    return pu::knn::make_layout(B, N1, N2).total;
}

int pu_knn_batch(const float *support, const float *query, int B, int N1, int N2, int K, int32_t *out_idx,
                 void *workspace, size_t workspace_bytes, pu_stream_t stream) {
    return pu::knn::knn_impl(support, query, B, N1, N2, K, out_idx, nullptr, workspace, workspace_bytes,
                             (cudaStream_t)stream);
}

int pu_knn_batch_dist(const float *support, const float *query, int B, int N1, int N2, int K, int32_t *out_idx,
                      float *out_dist, void *workspace, size_t workspace_bytes, pu_stream_t stream) {
    if (!out_dist) return PU_ERR_INVALID_ARG;
    return pu::knn::knn_impl(support, query, B, N1, N2, K, out_idx, out_dist, workspace, workspace_bytes,
                             (cudaStream_t)stream);
}

int pu_knn_self_interp(const float *cloud, int B, int N, int K, int n_sub, int32_t *out_neigh, int32_t *out_interp,
                       unsigned *out_unresolved, void *workspace, size_t workspace_bytes, pu_stream_t stream) {
    return pu::knn::knn_self_interp_impl(cloud, B, N, K, n_sub, out_neigh, out_interp, out_unresolved, workspace, workspace_bytes,
                                         (cudaStream_t)stream);
}

int pu_knn_read_stats(const void *workspace, unsigned long long *host_stats3, pu_stream_t stream) {
    if (!workspace || !host_stats3) return PU_ERR_INVALID_ARG;
    PU_CUDA_TRY(cudaMemcpyAsync(host_stats3, workspace, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                (cudaStream_t)stream));
    PU_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return PU_OK;
}

}  // extern "C"
