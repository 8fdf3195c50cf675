// knn.cu -- exact batched K-nearest-neighbour search for 3-D point clouds on sm_100a.
//
// Replaces the reference's host path  DataProcessing.knn_search (PointSegment/helper_tool.py:84-94)
//   -> nearest_neighbors.knn_batch (utils/nearest_neighbors/knn.pyx:71-109)
//   -> cpp_knn_batch_omp (utils/nearest_neighbors/knn_.cxx:104-135; nanoflann kd-tree per cloud).
// Not a port of the kd-tree: a GPU-native bucketed search.
//
//   build   (per call, per cloud)  Morton-order the support points (30-bit code, radix sort), cut the
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
#include <cub/device/device_radix_sort.cuh>
#include <cub/util_type.cuh>
#include <float.h>

#include "common.cuh"

namespace pu {
namespace knn {

constexpr int BS = 32;              // points per bucket  (= one warp-wide candidate tile)
constexpr int SBS = 32;             // buckets per super-bucket
constexpr int SEARCH_WARPS = 4;     // warps per CTA in the search kernel
constexpr unsigned long long KEY_INIT = 0x7F800000FFFFFFFFull;  // (+inf, id 0xFFFFFFFF)
constexpr size_t CUB_TEMP_RESERVE_BASE = 8u << 20;              // generous bound, checked at run time

struct Layout {
    size_t stats, bbox, keys_a, keys_b, vals_a, vals_b, sp, bk_lo, bk_hi, sb_lo, sb_hi;
    size_t qkeys_a, qkeys_b, qvals_a, qvals_b, sq, cub_temp, cub_bytes, total;
    int NB, NSB;
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
    L.stats = take(8 * sizeof(unsigned long long));
    L.bbox = take((size_t)B * 8 * sizeof(unsigned));
    L.keys_a = take(n1 * 8);
    L.keys_b = take(n1 * 8);
    L.vals_a = take(n1 * 4);
    L.vals_b = take(n1 * 4);
    L.sp = take(n1 * 16);
    L.bk_lo = take((size_t)B * L.NB * 16);
    L.bk_hi = take((size_t)B * L.NB * 16);
    L.sb_lo = take((size_t)B * L.NSB * 16);
    L.sb_hi = take((size_t)B * L.NSB * 16);
    L.qkeys_a = take(n2 * 8);
    L.qkeys_b = take(n2 * 8);
    L.qvals_a = take(n2 * 4);
    L.qvals_b = take(n2 * 4);
    L.sq = take(n2 * 16);
    L.cub_bytes = CUB_TEMP_RESERVE_BASE + (n1 > n2 ? n1 : n2);
    L.cub_temp = take(L.cub_bytes);
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
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            atomicMin(&bbox[b * 8 + c], f2ord(lo[c]));
            atomicMax(&bbox[b * 8 + 4 + c], f2ord(hi[c]));
        }
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

// key = (cloud << 30) | morton30(point quantised in the SUPPORT cloud's bounding box); val = global row
__global__ void __launch_bounds__(256) morton_kernel(const float *__restrict__ pts, int N, int B,
                                                     const unsigned *__restrict__ bbox,
                                                     unsigned long long *__restrict__ keys,
                                                     unsigned *__restrict__ vals) {
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
    keys[g] = ((unsigned long long)b << 30) | m;
    vals[g] = (unsigned)g;
}

// sorted[g] = (x, y, z, bits(local index))
__global__ void __launch_bounds__(256) gather_sorted_kernel(const float *__restrict__ pts, int N, int B,
                                                            const unsigned *__restrict__ vals_sorted,
                                                            float4 *__restrict__ out) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)B * N) return;
    const unsigned src = vals_sorted[g];
    const int b = (int)(g / N);
    const float *p = pts + (size_t)src * 3;
    out[g] = make_float4(p[0], p[1], p[2], __int_as_float((int)(src - (unsigned)b * (unsigned)N)));
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

    // all 32 lanes sweep the `cnt` candidates of one bucket staged in shared memory
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
                      const unsigned long long *__restrict__ skeys,  // sorted support keys (NULL for self-query)
                      const unsigned long long *__restrict__ qkeys,  // sorted query keys   (NULL for self-query)
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
            const unsigned long long key = qkeys[(size_t)b * N2 + w * 32 + (nvalid - 1) / 2];
            const unsigned long long *sk = skeys + (size_t)b * N1;
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

template <int K>
static int launch_search(const Layout &L, char *ws, bool self, const unsigned long long *skeys,
                         const unsigned long long *qkeys, const float4 *sq, int B, int N1, int N2, int kout,
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

static int sort_pairs(void *temp, size_t temp_reserved, unsigned long long *ka, unsigned long long *kb, unsigned *va,
                      unsigned *vb, size_t n, int end_bit, cudaStream_t st, unsigned long long **k_sorted,
                      unsigned **v_sorted) {
    cub::DoubleBuffer<unsigned long long> dk(ka, kb);
    cub::DoubleBuffer<unsigned> dv(va, vb);
    size_t need = 0;
    PU_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)n, 0, end_bit, st));
    if (need > temp_reserved) return PU_ERR_WORKSPACE;
    PU_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, need, dk, dv, (int)n, 0, end_bit, st));
    count_launch(3);
    *k_sorted = dk.Current();
    *v_sorted = dv.Current();
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
    int batch_bits = 0;
    while ((1 << batch_bits) < B) ++batch_bits;
    const int end_bit = 30 + batch_bits;
    const size_t n1 = (size_t)B * N1, n2 = (size_t)B * N2;
    unsigned *bbox = (unsigned *)(ws + L.bbox);

    PU_CUDA_TRY(cudaMemsetAsync(ws + L.stats, 0, 8 * sizeof(unsigned long long), st));
    bbox_init_kernel<<<ceil_div(B * 8, 128), 128, 0, st>>>(bbox, B);
    PU_LAUNCH_CHECK();
    {
        dim3 grid(min(ceil_div(N1, 256), 4 * kNumSMs), B);
        bbox_kernel<<<grid, 256, 0, st>>>(support, N1, bbox);
        PU_LAUNCH_CHECK();
    }
    morton_kernel<<<ceil_div(n1, 256), 256, 0, st>>>(support, N1, B, bbox, (unsigned long long *)(ws + L.keys_a),
                                                     (unsigned *)(ws + L.vals_a));
    PU_LAUNCH_CHECK();
    unsigned long long *skeys = nullptr, *qkeys = nullptr;
    unsigned *svals = nullptr, *qvals = nullptr;
    int rc = sort_pairs(ws + L.cub_temp, L.cub_bytes, (unsigned long long *)(ws + L.keys_a),
                        (unsigned long long *)(ws + L.keys_b), (unsigned *)(ws + L.vals_a), (unsigned *)(ws + L.vals_b),
                        n1, end_bit, st, &skeys, &svals);
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

    const float4 *sq = (const float4 *)(ws + L.sp);
    if (!self) {
        morton_kernel<<<ceil_div(n2, 256), 256, 0, st>>>(query, N2, B, bbox, (unsigned long long *)(ws + L.qkeys_a),
                                                         (unsigned *)(ws + L.qvals_a));
        PU_LAUNCH_CHECK();
        rc = sort_pairs(ws + L.cub_temp, L.cub_bytes, (unsigned long long *)(ws + L.qkeys_a),
                        (unsigned long long *)(ws + L.qkeys_b), (unsigned *)(ws + L.qvals_a),
                        (unsigned *)(ws + L.qvals_b), n2, end_bit, st, &qkeys, &qvals);
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

int pu_knn_read_stats(const void *workspace, unsigned long long *host_stats3, pu_stream_t stream) {
    if (!workspace || !host_stats3) return PU_ERR_INVALID_ARG;
    PU_CUDA_TRY(cudaMemcpyAsync(host_stats3, workspace, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                (cudaStream_t)stream));
    PU_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return PU_OK;
}

}  // extern "C"
