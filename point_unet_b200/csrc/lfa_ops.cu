// lfa_ops.cu -- the HBM-bound building blocks of the RandLA-Net local-feature-aggregation stack.
//
//   gather_rows        Network.gather_neighbour   (PointSegment/RandLANet.py:377-386)  tf.batch_gather
//                      Network.nearest_interpolation (:362-375)                        (K = 1)
//   locse              Network.relative_pos_encoding (:337-343)
//   random_sample      Network.random_sample (:345-360)  gather + reduce_max over K, fused
//   inverse lists      (no reference counterpart: TF back-propagates gathers with unsorted_segment_sum,
//                      i.e. atomics; here every gather gradient is a per-point SEGMENTED SUM over a
//                      precomputed inverse neighbour list -- scatter-free and bit-deterministic)
//
// All tensors channels-last fp32; "rows" of d channels are moved as 128-bit vectors when d % 4 == 0 and the
// row strides allow it.  Outputs take a row stride (ld) so a kernel can write straight into one half of a
// concat buffer (tf.concat at RandLANet.py:328,333,138 never needs its own copy).
#include <float.h>

#include "common.cuh"

namespace pu {
namespace lfa {

// ---------------------------------------------------------------------------------------------
// out[b, r, 0:d] = src[b, idx[b, r], 0:d]        r in [0, R)   (R = M*K rows per cloud)
// 128-bit path: grid (x, B); a thread owns ONE 16-byte column chunk (fixed for its lifetime: no div/mod in the loop) and
// walks rows with 4 independent (index -> row -> streaming store) chains in flight.
template <int UNROLL>
__global__ void __launch_bounds__(256) gather_rows_v4_kernel(const float *__restrict__ src, int ld_src, int n_src,
                                                             const int32_t *__restrict__ idx, int R,
                                                             float *__restrict__ dst, int ld_dst, int cpr) {
    const int b = blockIdx.y;
    const int c = (threadIdx.x % cpr) * 4, rl = threadIdx.x / cpr, rpb = 256 / cpr;
    if (rl >= rpb) return;
    const float *sb = src + (size_t)b * n_src * ld_src + c;
    const int32_t *ib = idx + (size_t)b * R;
    float *db = dst + (size_t)b * R * ld_dst + c;
    const int step = gridDim.x * rpb;
    int r = blockIdx.x * rpb + rl;
    for (; r + (UNROLL - 1) * step < R; r += UNROLL * step) {
        int j[UNROLL];
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) j[u] = ib[r + u * step];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = *reinterpret_cast<const float4 *>(sb + (size_t)j[u] * ld_src);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) st_stream_f4(reinterpret_cast<float4 *>(db + (size_t)(r + u * step) * ld_dst), v[u]);
    }
    for (; r < R; r += step)
        st_stream_f4(reinterpret_cast<float4 *>(db + (size_t)r * ld_dst),
                     *reinterpret_cast<const float4 *>(sb + (size_t)ib[r] * ld_src));
}

template <int VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ src, int ld_src, int n_src,
                                                          const int32_t *__restrict__ idx, long long R, int B,
                                                          float *__restrict__ dst, int ld_dst, int d) {
    const int cpr = d / VEC;  // chunks per row
    const long long total = (long long)B * R * cpr;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / cpr;
        const int c = (int)(t - row * cpr) * VEC;
        const int b = (int)(row / R);
        const int j = idx[row];
        const float *s = src + ((size_t)b * n_src + j) * ld_src + c;
        float *o = dst + (size_t)row * ld_dst + c;
        if (VEC == 4) {
            st_stream_f4(reinterpret_cast<float4 *>(o), *reinterpret_cast<const float4 *>(s));
        } else {
            *o = *s;
        }
    }
}

// grad_src[b, j, 0:d] = sum over e in [off[b*n+j], off[b*n+j+1]) of grad_out[perm[e], 0:d]
// (perm holds GLOBAL row numbers b*R + r, ascending inside a segment => fixed summation order)
// 128-bit path: a thread owns one 16-byte column chunk; the segment is walked 4 edges at a time (4 independent
// perm -> row loads in flight), added in the fixed ascending order.
__global__ void __launch_bounds__(256) segment_sum_v4_kernel(const float *__restrict__ grad_out, int ld_go,
                                                             const int32_t *__restrict__ off,
                                                             const int32_t *__restrict__ perm, long long n_targets,
                                                             float *__restrict__ grad_src, int ld_gs, int cpr,
                                                             int accumulate) {
    const int c = (threadIdx.x % cpr) * 4, rl = threadIdx.x / cpr, rpb = 256 / cpr;
    if (rl >= rpb) return;
    const float *g = grad_out + c;
    for (long long j = (long long)blockIdx.x * rpb + rl; j < n_targets; j += (long long)gridDim.x * rpb) {
        const int e0 = off[j], e1 = off[j + 1];
        float *o = grad_src + (size_t)j * ld_gs + c;
        float4 acc = accumulate ? *reinterpret_cast<float4 *>(o) : make_float4(0.f, 0.f, 0.f, 0.f);
        int e = e0;
        for (; e + 4 <= e1; e += 4) {
            const int p0 = perm[e], p1 = perm[e + 1], p2 = perm[e + 2], p3 = perm[e + 3];
            const float4 a = ld_stream_f4(reinterpret_cast<const float4 *>(g + (size_t)p0 * ld_go));
            const float4 b = ld_stream_f4(reinterpret_cast<const float4 *>(g + (size_t)p1 * ld_go));
            const float4 cc = ld_stream_f4(reinterpret_cast<const float4 *>(g + (size_t)p2 * ld_go));
            const float4 dd = ld_stream_f4(reinterpret_cast<const float4 *>(g + (size_t)p3 * ld_go));
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
            acc.x += cc.x; acc.y += cc.y; acc.z += cc.z; acc.w += cc.w;
            acc.x += dd.x; acc.y += dd.y; acc.z += dd.z; acc.w += dd.w;
        }
        for (; e < e1; ++e) {
            const float4 a = ld_stream_f4(reinterpret_cast<const float4 *>(g + (size_t)perm[e] * ld_go));
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        }
        *reinterpret_cast<float4 *>(o) = acc;
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) segment_sum_kernel(const float *__restrict__ grad_out, int ld_go,
                                                          const int32_t *__restrict__ off,
                                                          const int32_t *__restrict__ perm, long long n_targets,
                                                          float *__restrict__ grad_src, int ld_gs, int d,
                                                          int accumulate) {
    const int cpr = d / VEC;
    const long long total = n_targets * cpr;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long j = t / cpr;
        const int c = (int)(t - j * cpr) * VEC;
        const int e0 = off[j], e1 = off[j + 1];
        float *o = grad_src + (size_t)j * ld_gs + c;
        float acc = accumulate ? *o : 0.f;
        for (int e = e0; e < e1; ++e) acc += grad_out[(size_t)perm[e] * ld_go + c];
        *o = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// inverse neighbour lists.  off[t] .. off[t+1] delimits the rows (global numbers b*R + r) whose index points at target
// t = b*n + idx[row]; inside a segment the rows are ASCENDING, so every segmented sum adds in one fixed order.
// Built as a counting sort -- the lists are short (16 on average, in-degree of a K-NN graph), a general radix sort of
// 11.5 M (key, row) pairs was three full passes over them:
//   1. cnt[t]++                 integer atomics (the counts do not depend on their order)
//   2. off = exclusive scan(cnt)
//   3. tmp[off[t] + slot] = row slot handed out by atomicSub on cnt[t]: the order inside a segment is arbitrary here ...
//   4. perm[off[t] + rank] = row  ... and made canonical by ranking every row within its segment (rows are distinct).
//      Neighbouring threads rank rows of the same segment, so its 64 bytes are read once and broadcast.  Segments longer
//      than INV_LONG (clouds with thousands of coincident points) are ranked by a whole CTA instead of one thread per row.
constexpr int INV_LONG = 1024;

// exclusive prefix sum of the per-target counts (in-tree; three launches):
//   scan_block_sums   CTA i: sum of its 2048 counts -> bsum[i]
//   scan_block_offs   one CTA: exclusive scan of bsum in place (chunks of 1024 with a running carry)
//   scan_apply        CTA i: exclusive scan of its 2048 counts (8 per thread, warp shuffles) + bsum[i]
constexpr int SCAN_THREADS = 256, SCAN_IPT = 8, SCAN_TILE = SCAN_THREADS * SCAN_IPT;

__device__ __forceinline__ int block_exclusive_scan_256(int v, int *s_warp, int &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < SCAN_THREADS / 32; ++i) {
        const int t = s_warp[i];
        if (i < w) base += t;
        tot += t;
    }
    total = tot;
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const int32_t *__restrict__ cnt, long long n,
                                                                       int32_t *__restrict__ bsum) {
    __shared__ int s_warp[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    int v = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        const long long j = base + (long long)i * SCAN_THREADS + threadIdx.x;  // coalesced
        if (j < n) v += cnt[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int i = 0; i < SCAN_THREADS / 32; ++i) t += s_warp[i];
        bsum[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_offs_kernel(int32_t *__restrict__ bsum, int nblocks) {
    __shared__ int s_warp[SCAN_THREADS / 32];
    int carry = 0;
    for (int c0 = 0; c0 < nblocks; c0 += SCAN_THREADS) {
        const int i = c0 + threadIdx.x;
        const int v = i < nblocks ? bsum[i] : 0;
        int total;
        const int ex = block_exclusive_scan_256(v, s_warp, total);
        if (i < nblocks) bsum[i] = carry + ex;
        carry += total;
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int32_t *__restrict__ cnt, long long n,
                                                                  const int32_t *__restrict__ bsum, int32_t *__restrict__ out) {
    __shared__ int s_warp[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_IPT;  // 8 consecutive per thread
    int v[SCAN_IPT], sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        v[i] = base + i < n ? cnt[base + i] : 0;
        sum += v[i];
    }
    int total;
    int run = block_exclusive_scan_256(sum, s_warp, total) + bsum[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

__global__ void __launch_bounds__(256) inv_count_kernel(const int32_t *__restrict__ idx, long long R, int B, int n,
                                                        int32_t *__restrict__ cnt) {
    const unsigned total = (unsigned)((long long)B * R), Ru = (unsigned)R;  // < 2^31 (checked by the launcher): 32-bit division
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const unsigned b = t / Ru;
        atomicAdd(&cnt[(size_t)b * n + idx[t]], 1);
    }
}
__global__ void __launch_bounds__(256) inv_fill_kernel(const int32_t *__restrict__ idx, long long R, int B, int n,
                                                       const int32_t *__restrict__ off, int32_t *__restrict__ cnt,
                                                       int32_t *__restrict__ tmp) {
    const unsigned total = (unsigned)((long long)B * R), Ru = (unsigned)R;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const unsigned b = t / Ru;
        const size_t key = (size_t)b * n + idx[t];
        const int slot = atomicSub(&cnt[key], 1) - 1;
        tmp[off[key] + slot] = (int32_t)t;
    }
}
__global__ void __launch_bounds__(256) inv_rank_kernel(const int32_t *__restrict__ idx, long long R, int n,
                                                       const int32_t *__restrict__ off, const int32_t *__restrict__ tmp,
                                                       long long total, int32_t *__restrict__ perm) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int row = tmp[e];
        const size_t key = (size_t)((unsigned)row / (unsigned)R) * n + idx[row];
        const int e0 = off[key], e1 = off[key + 1];
        if (e1 - e0 > INV_LONG) continue;  // inv_rank_long_kernel
        int rank = 0;
        for (int i = e0; i < e1; ++i) rank += tmp[i] < row;
        perm[e0 + rank] = row;
    }
}
// long segments: each CTA looks at 256 targets at a time, collects the long ones and ranks them with all its threads
__global__ void __launch_bounds__(256) inv_rank_long_kernel(const int32_t *__restrict__ off, const int32_t *__restrict__ tmp,
                                                            long long n_targets, int32_t *__restrict__ perm) {
    __shared__ int s_list[256];
    __shared__ int s_n;
    for (long long base = (long long)blockIdx.x * 256; base < n_targets; base += (long long)gridDim.x * 256) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const long long t = base + threadIdx.x;
        if (t < n_targets && off[t + 1] - off[t] > INV_LONG) s_list[atomicAdd(&s_n, 1)] = (int)threadIdx.x;
        __syncthreads();
        const int nl = s_n;
        for (int q = 0; q < nl; ++q) {
            const long long tt = base + s_list[q];
            const int e0 = off[tt], e1 = off[tt + 1];
            for (int i = e0 + threadIdx.x; i < e1; i += 256) {
                const int row = tmp[i];
                int rank = 0;
                for (int j = e0; j < e1; ++j) rank += tmp[j] < row;
                perm[e0 + rank] = row;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// LocSE: out[b,n,k,:] = [ |p - q|, p - q (3), p (3), q (3) ]  with p = xyz[b,n], q = xyz[b, idx[b,n,k]]
// grid (tiles of 256 (n,k) rows inside a cloud, B): the cloud comes from blockIdx.y and the point from a 32-bit shift (K a
// power of two) or 32-bit division -- the first version spent ~150 of its ~250 instructions per row on two 64-bit
// divisions (row / K, point / N).  A tile = 2560 contiguous floats, staged in shared memory and stored as float4
// (the rows of a cloud start at a multiple of 4 floats whenever N*K is even; otherwise the launcher asks for scalar stores).
__global__ void __launch_bounds__(256) locse_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ idx,
                                                    int N, int K, int shiftK, unsigned rows_per_cloud, int vec_ok,
                                                    float *__restrict__ out) {
    __shared__ __align__(16) float s_out[256 * 10];
    const unsigned r0 = blockIdx.x * 256u, r = r0 + threadIdx.x;
    const size_t cloud_row0 = (size_t)blockIdx.y * rows_per_cloud;
    const float *xb = xyz + (size_t)blockIdx.y * N * 3;
    if (r < rows_per_cloud) {
        const unsigned n = shiftK >= 0 ? (r >> shiftK) : (r / (unsigned)K);
        const int j = idx[cloud_row0 + r];
        const float *p = xb + (size_t)n * 3;
        const float *q = xb + (size_t)j * 3;
        const float px = p[0], py = p[1], pz = p[2], qx = q[0], qy = q[1], qz = q[2];
        const float rx = px - qx, ry = py - qy, rz = pz - qz;
        float *o = s_out + threadIdx.x * 10;
        o[0] = sqrtf(rx * rx + ry * ry + rz * rz);
        o[1] = rx; o[2] = ry; o[3] = rz;
        o[4] = px; o[5] = py; o[6] = pz;
        o[7] = qx; o[8] = qy; o[9] = qz;
    }
    __syncthreads();
    const unsigned nrows = min(256u, rows_per_cloud - r0);
    const int nvec = vec_ok ? (int)(nrows * 10 / 4) : 0;
    float *ob = out + (cloud_row0 + r0) * 10;
    float4 *o4 = reinterpret_cast<float4 *>(ob);
    for (int v = threadIdx.x; v < nvec; v += 256) st_stream_f4(o4 + v, reinterpret_cast<float4 *>(s_out)[v]);
    for (int t = nvec * 4 + threadIdx.x; t < (int)nrows * 10; t += 256) ob[t] = s_out[t];
}

// ---------------------------------------------------------------------------------------------
// random_sample: out[b,m,c] = max_k feat[b, idx[b,m,k], c];  ties[b,m,c] = number of k attaining the max
template <int VEC>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float *__restrict__ feat, int ld_f, int n_src,
                                                          const int32_t *__restrict__ idx, int M, int K, int B,
                                                          float *__restrict__ out, int ld_o,
                                                          unsigned char *__restrict__ ties, int d) {
    const int cpr = d / VEC;
    const long long total = (long long)B * M * cpr;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long m = t / cpr;  // global output point b*M + m
        const int c = (int)(t - m * cpr) * VEC;
        const int b = (int)(m / M);
        const int32_t *ix = idx + (size_t)m * K;
        float best[VEC];
        int cnt[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { best[v] = -FLT_MAX; cnt[v] = 0; }
        for (int k = 0; k < K; ++k) {
            const float *s = feat + ((size_t)b * n_src + ix[k]) * ld_f + c;
            float x[VEC];
            if (VEC == 4) {
                const float4 v4 = *reinterpret_cast<const float4 *>(s);
                x[0] = v4.x; x[1 % VEC] = v4.y; x[2 % VEC] = v4.z; x[3 % VEC] = v4.w;
            } else {
                x[0] = *s;
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (x[v] > best[v]) { best[v] = x[v]; cnt[v] = 1; }
                else if (x[v] == best[v]) cnt[v]++;
            }
        }
        float *o = out + (size_t)m * ld_o + c;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            o[v] = best[v];
            if (ties) ties[(size_t)m * d + c + v] = (unsigned char)cnt[v];
        }
    }
}

// gradient of random_sample, scatter-free: every source point j walks its inverse list (edges e = m*K + k of
// pool_idx that reference j) and takes g[m,c] / ties[m,c] wherever its own value is the maximum
// (tf.reduce_max's gradient splits evenly among exact ties).
template <int VEC>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float *__restrict__ feat, int ld_f,
                                                          const float *__restrict__ out, int ld_o,
                                                          const unsigned char *__restrict__ ties,
                                                          const float *__restrict__ g_out, int ld_g,
                                                          const int32_t *__restrict__ off,
                                                          const int32_t *__restrict__ perm, long long n_targets,
                                                          int K, float *__restrict__ g_feat, int ld_gf, int d) {
    const int cpr = d / VEC;
    const long long total = n_targets * cpr;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long j = t / cpr;
        const int c = (int)(t - j * cpr) * VEC;
        float acc[VEC], x[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { acc[v] = 0.f; x[v] = feat[(size_t)j * ld_f + c + v]; }
        const int e0 = off[j], e1 = off[j + 1];
        for (int e = e0; e < e1; ++e) {
            const long long m = perm[e] / K;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float o = out[(size_t)m * ld_o + c + v];
                if (x[v] == o) acc[v] += g_out[(size_t)m * ld_g + c + v] / (float)ties[(size_t)m * d + c + v];
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) g_feat[(size_t)j * ld_gf + c + v] = acc[v];
    }
}

// ---------------------------------------------------------------------------------------------
// point2prod (PointSegment/testPancreas.py:71-85, testBraTS.py:83-101): scatter per-point class probabilities into
// a dense volume.  The reference loops `volume[z][x][y] = prob[i]` on the host and then moves axis 1 <-> 2; here the
// result is written directly in the final [Z, Y, X, C] layout.  Sequential semantics (the LAST point writing a voxel
// wins) are kept deterministically: pass 1 records the highest point index per voxel, pass 2 lets only that point write.
__global__ void __launch_bounds__(256) p2v_owner_kernel(const int32_t *__restrict__ xyz_origin,
                                                        const int32_t *__restrict__ point_idx, int n, int Z, int X, int Y,
                                                        int32_t *__restrict__ owner) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t *o = xyz_origin + (size_t)(point_idx ? point_idx[i] : i) * 3;
    const int x = o[0], y = o[1], z = o[2];
    if ((unsigned)x >= (unsigned)X || (unsigned)y >= (unsigned)Y || (unsigned)z >= (unsigned)Z) return;
    atomicMax(&owner[((size_t)z * Y + y) * X + x], i);
}
__global__ void __launch_bounds__(256) p2v_write_kernel(const float *__restrict__ probs, const int32_t *__restrict__ xyz_origin,
                                                        const int32_t *__restrict__ point_idx, int n, int C, int Z, int X,
                                                        int Y, const int32_t *__restrict__ owner, float *__restrict__ vol) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t *o = xyz_origin + (size_t)(point_idx ? point_idx[i] : i) * 3;
    const int x = o[0], y = o[1], z = o[2];
    if ((unsigned)x >= (unsigned)X || (unsigned)y >= (unsigned)Y || (unsigned)z >= (unsigned)Z) return;
    const size_t v = ((size_t)z * Y + y) * X + x;
    if (owner[v] != i) return;
    for (int c = 0; c < C; ++c) vol[v * C + c] = probs[(size_t)i * C + c];
}

// label volume (utils/genSegmentationPancreas.py:67-77, genSegmentationBraTS.py:67-78): seg = argmax(prob volume, -1) as
// uint8 (first maximum wins, like np.argmax; untouched voxels hold all-zero probabilities -> label 0), BraTS maps 3 -> 4.
// Written straight from the per-point probabilities by the voxel's owner: the dense fp32 probability volume
// (0.25-1 GB per case in the reference) is never materialised.
__global__ void __launch_bounds__(256) p2v_label_kernel(const float *__restrict__ probs, const int32_t *__restrict__ xyz_origin,
                                                        const int32_t *__restrict__ point_idx, int n, int C, int Z, int X,
                                                        int Y, const int32_t *__restrict__ owner, int remap_from, int remap_to,
                                                        unsigned char *__restrict__ labels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t *o = xyz_origin + (size_t)(point_idx ? point_idx[i] : i) * 3;
    const int x = o[0], y = o[1], z = o[2];
    if ((unsigned)x >= (unsigned)X || (unsigned)y >= (unsigned)Y || (unsigned)z >= (unsigned)Z) return;
    const size_t v = ((size_t)z * Y + y) * X + x;
    if (owner[v] != i) return;
    const float *p = probs + (size_t)i * C;
    int best = 0;
    float bv = p[0];
    for (int c = 1; c < C; ++c)
        if (p[c] > bv) { bv = p[c]; best = c; }
    labels[v] = (unsigned char)(best == remap_from ? remap_to : best);
}
// argmax over the last axis of an existing dense volume [nvox, C] -> uint8 [nvox]
__global__ void __launch_bounds__(256) volume_argmax_kernel(const float *__restrict__ vol, long long nvox, int C, int remap_from,
                                                            int remap_to, unsigned char *__restrict__ labels) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
        const float *p = vol + (size_t)v * C;
        int best = 0;
        float bv = p[0];
        for (int c = 1; c < C; ++c)
            if (p[c] > bv) { bv = p[c]; best = c; }
        labels[v] = (unsigned char)(best == remap_from ? remap_to : best);
    }
}

// 128-bit max-pool forward for K = 16: a thread owns one 16-byte column chunk of one output point, loads the 16
// neighbour ids as four int4 and keeps all 16 row loads in flight before reducing.
__global__ void __launch_bounds__(256) maxpool_fwd_v4_kernel(const float *__restrict__ feat, int ld_f, int n_src,
                                                             const int32_t *__restrict__ idx, int M,
                                                             float *__restrict__ out, int ld_o,
                                                             unsigned char *__restrict__ ties, int d, int cpr) {
    constexpr int K = 16;
    const int b = blockIdx.y;
    const int c = (threadIdx.x % cpr) * 4, rl = threadIdx.x / cpr, rpb = 256 / cpr;
    if (rl >= rpb) return;
    const float *fb = feat + (size_t)b * n_src * ld_f + c;
    for (int m = blockIdx.x * rpb + rl; m < M; m += gridDim.x * rpb) {
        const size_t gm = (size_t)b * M + m;
        const int4 *ip = reinterpret_cast<const int4 *>(idx + gm * K);
        int j[K];
#pragma unroll
        for (int q = 0; q < K / 4; ++q) {
            const int4 t = ip[q];
            j[4 * q] = t.x; j[4 * q + 1] = t.y; j[4 * q + 2] = t.z; j[4 * q + 3] = t.w;
        }
        float4 v[K];
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = *reinterpret_cast<const float4 *>(fb + (size_t)j[k] * ld_f);
        float4 best = v[0];
#pragma unroll
        for (int k = 1; k < K; ++k) {
            best.x = fmaxf(best.x, v[k].x); best.y = fmaxf(best.y, v[k].y);
            best.z = fmaxf(best.z, v[k].z); best.w = fmaxf(best.w, v[k].w);
        }
        st_stream_f4(reinterpret_cast<float4 *>(out + gm * ld_o + c), best);
        if (ties) {
            int cx = 0, cy = 0, cz = 0, cw = 0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                cx += v[k].x == best.x; cy += v[k].y == best.y; cz += v[k].z == best.z; cw += v[k].w == best.w;
            }
            *reinterpret_cast<uchar4 *>(ties + gm * d + c) = make_uchar4((unsigned char)cx, (unsigned char)cy, (unsigned char)cz, (unsigned char)cw);
        }
    }
}

// 128-bit max-pool backward: fixed column chunk per thread, the (short) inverse list walked two edges at a time
__global__ void __launch_bounds__(256) maxpool_bwd_v4_kernel(const float *__restrict__ feat, int ld_f,
                                                             const float *__restrict__ out, int ld_o,
                                                             const unsigned char *__restrict__ ties,
                                                             const float *__restrict__ g_out, int ld_g,
                                                             const int32_t *__restrict__ off,
                                                             const int32_t *__restrict__ perm, long long n_targets,
                                                             int K, float *__restrict__ g_feat, int ld_gf, int d, int cpr) {
    const int c = (threadIdx.x % cpr) * 4, rl = threadIdx.x / cpr, rpb = 256 / cpr;
    if (rl >= rpb) return;
    for (long long j = (long long)blockIdx.x * rpb + rl; j < n_targets; j += (long long)gridDim.x * rpb) {
        const float4 x = *reinterpret_cast<const float4 *>(feat + (size_t)j * ld_f + c);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int e0 = off[j], e1 = off[j + 1];
        int e = e0;
        auto add = [&](const float4 &o, const float4 &g, const uchar4 &t) {
            if (x.x == o.x) acc.x += g.x / (float)t.x;
            if (x.y == o.y) acc.y += g.y / (float)t.y;
            if (x.z == o.z) acc.z += g.z / (float)t.z;
            if (x.w == o.w) acc.w += g.w / (float)t.w;
        };
        for (; e + 2 <= e1; e += 2) {
            const long long m0 = perm[e] / K, m1 = perm[e + 1] / K;
            const float4 o0 = *reinterpret_cast<const float4 *>(out + (size_t)m0 * ld_o + c);
            const float4 o1 = *reinterpret_cast<const float4 *>(out + (size_t)m1 * ld_o + c);
            const float4 g0 = *reinterpret_cast<const float4 *>(g_out + (size_t)m0 * ld_g + c);
            const float4 g1 = *reinterpret_cast<const float4 *>(g_out + (size_t)m1 * ld_g + c);
            const uchar4 t0 = *reinterpret_cast<const uchar4 *>(ties + (size_t)m0 * d + c);
            const uchar4 t1 = *reinterpret_cast<const uchar4 *>(ties + (size_t)m1 * d + c);
            add(o0, g0, t0);
            add(o1, g1, t1);
        }
        for (; e < e1; ++e) {
            const long long m0 = perm[e] / K;
            add(*reinterpret_cast<const float4 *>(out + (size_t)m0 * ld_o + c),
                *reinterpret_cast<const float4 *>(g_out + (size_t)m0 * ld_g + c),
                *reinterpret_cast<const uchar4 *>(ties + (size_t)m0 * d + c));
        }
        st_stream_f4(reinterpret_cast<float4 *>(g_feat + (size_t)j * ld_gf + c), acc);
    }
}

static inline int grid_for(long long total, int block = 256) {
    long long g = (total + block - 1) / block;
    const long long cap = (long long)kNumSMs * 32;  // grid-stride loops; cap at 32 CTAs per SM
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
static inline bool vec4_ok(const void *p, int ld) { return (((uintptr_t)p) & 15) == 0 && (ld & 3) == 0; }

}  // namespace lfa
}  // namespace pu

using namespace pu;
using namespace pu::lfa;

extern "C" {

int pu_gather_rows_fwd(const float *src, int ld_src, int n_src, const int32_t *idx, long long rows_per_cloud, int B,
                       float *dst, int ld_dst, int d, pu_stream_t stream) {
    if (!src || !idx || !dst || B < 0 || rows_per_cloud < 0 || d < 1 || ld_src < d || ld_dst < d || n_src < 1)
        return PU_ERR_INVALID_ARG;
    if (B == 0 || rows_per_cloud == 0) return PU_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if ((d & 3) == 0 && d <= 1024 && rows_per_cloud < (1ll << 31) && vec4_ok(src, ld_src) && vec4_ok(dst, ld_dst)) {
        const int cpr = d / 4, rpb = 256 / cpr;
        long long gx = (rows_per_cloud + (long long)rpb * 4 - 1) / ((long long)rpb * 4);  // ~4 rows per thread
        const long long cap = (long long)kNumSMs * 16 / (B > 0 ? B : 1) + 1;
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        dim3 grid((unsigned)gx, B);
        gather_rows_v4_kernel<4><<<grid, 256, 0, st>>>(src, ld_src, n_src, idx, (int)rows_per_cloud, dst, ld_dst, cpr);
    } else {
        gather_rows_kernel<1><<<grid_for((long long)B * rows_per_cloud * d), 256, 0, st>>>(
            src, ld_src, n_src, idx, rows_per_cloud, B, dst, ld_dst, d);
    }
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_segment_sum(const float *grad_out, int ld_go, const int32_t *offsets, const int32_t *perm, long long n_targets,
                   float *grad_src, int ld_gs, int d, int accumulate, pu_stream_t stream) {
    if (!grad_out || !offsets || !perm || !grad_src || n_targets < 0 || d < 1 || ld_go < d || ld_gs < d)
        return PU_ERR_INVALID_ARG;
    if (n_targets == 0) return PU_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if ((d & 3) == 0 && d <= 1024 && vec4_ok(grad_out, ld_go) && vec4_ok(grad_src, ld_gs)) {
        const int cpr = d / 4, rpb = 256 / cpr;
        long long gx = (n_targets + rpb - 1) / rpb;
        const long long cap = (long long)kNumSMs * 16;
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        segment_sum_v4_kernel<<<(unsigned)gx, 256, 0, st>>>(grad_out, ld_go, offsets, perm, n_targets, grad_src, ld_gs, cpr,
                                                           accumulate);
    } else {
        segment_sum_kernel<1><<<grid_for(n_targets * d), 256, 0, st>>>(grad_out, ld_go, offsets, perm, n_targets,
                                                                    grad_src, ld_gs, d, accumulate);
    }
    PU_LAUNCH_CHECK();
    return PU_OK;
}

size_t pu_inverse_workspace_bytes(int B, long long rows_per_cloud) {
    if (B <= 0 || rows_per_cloud < 0) return 0;
    const size_t n = (size_t)B * rows_per_cloud;
    // unsorted lists [n] + counters [<= n + 1 targets ... sized below from n_src at call time] + scan scratch
    return 4 * align_up(n * 4, 256) + (8u << 20) + n;
}

int pu_build_inverse(const int32_t *idx, long long rows_per_cloud, int B, int n_src, int32_t *offsets, int32_t *perm,
                     void *workspace, size_t workspace_bytes, pu_stream_t stream) {
    if (!idx || !offsets || !perm || B < 0 || rows_per_cloud < 0 || n_src < 1) return PU_ERR_INVALID_ARG;
    const long long total = (long long)B * rows_per_cloud, n_targets = (long long)B * n_src;
    if (total >= (1ll << 31) || n_targets >= (1ll << 31)) return PU_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (!workspace || workspace_bytes < pu_inverse_workspace_bytes(B, rows_per_cloud)) return PU_ERR_WORKSPACE;
    if (total == 0) {
        PU_CUDA_TRY(cudaMemsetAsync(offsets, 0, (size_t)(n_targets + 1) * 4, st));
        return PU_OK;
    }
    char *ws = (char *)workspace;
    const size_t seg_tmp = align_up((size_t)total * 4, 256), seg_cnt = align_up((size_t)(n_targets + 1) * 4, 256);
    if (seg_tmp + seg_cnt + (1u << 20) > workspace_bytes) return PU_ERR_WORKSPACE;  // many more targets than rows
    int32_t *tmp = (int32_t *)ws, *cnt = (int32_t *)(ws + seg_tmp);
    void *temp = ws + seg_tmp + seg_cnt;
    const size_t temp_reserved = workspace_bytes - seg_tmp - seg_cnt;
    PU_CUDA_TRY(cudaMemsetAsync(cnt, 0, (size_t)(n_targets + 1) * 4, st));
    inv_count_kernel<<<grid_for(total), 256, 0, st>>>(idx, rows_per_cloud, B, n_src, cnt);
    PU_LAUNCH_CHECK();
    {   // offsets = exclusive scan of the counts (n_targets + 1 entries: the last one is the total)
        const long long n = n_targets + 1;
        const int nblocks = ceil_div(n, SCAN_TILE);
        if ((size_t)nblocks * sizeof(int32_t) > temp_reserved) return PU_ERR_WORKSPACE;
        int32_t *bsum = (int32_t *)temp;
        scan_block_sums_kernel<<<nblocks, SCAN_THREADS, 0, st>>>(cnt, n, bsum);
        PU_LAUNCH_CHECK();
        scan_block_offs_kernel<<<1, SCAN_THREADS, 0, st>>>(bsum, nblocks);
        PU_LAUNCH_CHECK();
        scan_apply_kernel<<<nblocks, SCAN_THREADS, 0, st>>>(cnt, n, bsum, offsets);
        PU_LAUNCH_CHECK();
    }
    inv_fill_kernel<<<grid_for(total), 256, 0, st>>>(idx, rows_per_cloud, B, n_src, offsets, cnt, tmp);
    PU_LAUNCH_CHECK();
    inv_rank_kernel<<<grid_for(total), 256, 0, st>>>(idx, rows_per_cloud, n_src, offsets, tmp, total, perm);
    PU_LAUNCH_CHECK();
    inv_rank_long_kernel<<<grid_for(n_targets), 256, 0, st>>>(offsets, tmp, n_targets, perm);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_relative_pos_encoding_fwd(const float *xyz, const int32_t *idx, int B, int N, int K, float *out,
                                 pu_stream_t stream) {
    if (!xyz || !idx || !out || B < 0 || N < 0 || K < 1) return PU_ERR_INVALID_ARG;
    const long long rpc = (long long)N * K;
    if (B == 0 || rpc == 0) return PU_OK;
    if ((((uintptr_t)out) & 15) != 0) return PU_ERR_INVALID_ARG;
    if (rpc >= (1ll << 31) || B > 65535) return PU_ERR_UNSUPPORTED;
    const int vec_ok = (B == 1 || (rpc & 1) == 0) ? 1 : 0;  // 128-bit stores need every cloud to start at a multiple of 4 floats
    int shiftK = -1;
    if ((K & (K - 1)) == 0) { shiftK = 0; while ((1 << shiftK) < K) ++shiftK; }
    dim3 grid((unsigned)ceil_div(rpc, 256), (unsigned)B);
    locse_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, idx, N, K, shiftK, (unsigned)rpc, vec_ok, out);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_random_sample_fwd(const float *feat, int ld_f, int n_src, const int32_t *pool_idx, int B, int M, int K,
                         float *out, int ld_o, unsigned char *ties, int d, pu_stream_t stream) {
    if (!feat || !pool_idx || !out || B < 0 || M < 0 || K < 1 || K > 255 || d < 1 || ld_f < d || ld_o < d || n_src < 1)
        return PU_ERR_INVALID_ARG;
    if (B == 0 || M == 0) return PU_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 16 && (d & 3) == 0 && d <= 1024 && vec4_ok(feat, ld_f) && vec4_ok(out, ld_o) &&
        ((((uintptr_t)pool_idx) & 15) == 0) && (!ties || ((((uintptr_t)ties) & 3) == 0))) {
        const int cpr = d / 4, rpb = 256 / cpr;
        long long gx = ((long long)M + rpb - 1) / rpb;
        const long long cap = (long long)kNumSMs * 16 / (B > 0 ? B : 1) + 1;
        if (gx > cap) gx = cap;
        dim3 grid((unsigned)gx, B);
        maxpool_fwd_v4_kernel<<<grid, 256, 0, st>>>(feat, ld_f, n_src, pool_idx, M, out, ld_o, ties, d, cpr);
    } else if ((d & 3) == 0 && vec4_ok(feat, ld_f) && vec4_ok(out, ld_o)) {
        maxpool_fwd_kernel<4><<<grid_for((long long)B * M * (d / 4)), 256, 0, st>>>(feat, ld_f, n_src, pool_idx, M, K, B,
                                                                                  out, ld_o, ties, d);
    } else {
        maxpool_fwd_kernel<1><<<grid_for((long long)B * M * d), 256, 0, st>>>(feat, ld_f, n_src, pool_idx, M, K, B, out,
                                                                            ld_o, ties, d);
    }
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_random_sample_bwd(const float *feat, int ld_f, const float *out, int ld_o, const unsigned char *ties,
                         const float *g_out, int ld_g, const int32_t *offsets, const int32_t *perm,
                         long long n_targets, int K, float *g_feat, int ld_gf, int d, pu_stream_t stream) {
    if (!feat || !out || !ties || !g_out || !offsets || !perm || !g_feat || n_targets < 0 || K < 1 || d < 1)
        return PU_ERR_INVALID_ARG;
    if (n_targets == 0) return PU_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if ((d & 3) == 0 && d <= 1024 && vec4_ok(feat, ld_f) && vec4_ok(out, ld_o) && vec4_ok(g_out, ld_g) &&
        vec4_ok(g_feat, ld_gf) && ((((uintptr_t)ties) & 3) == 0)) {
        const int cpr = d / 4, rpb = 256 / cpr;
        long long gx = (n_targets + rpb - 1) / rpb;
        const long long cap = (long long)kNumSMs * 16;
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        maxpool_bwd_v4_kernel<<<(unsigned)gx, 256, 0, st>>>(feat, ld_f, out, ld_o, ties, g_out, ld_g, offsets, perm, n_targets,
                                                           K, g_feat, ld_gf, d, cpr);
    } else if ((d & 3) == 0) {
        maxpool_bwd_kernel<4><<<grid_for(n_targets * (d / 4)), 256, 0, st>>>(feat, ld_f, out, ld_o, ties, g_out, ld_g,
                                                                           offsets, perm, n_targets, K, g_feat, ld_gf, d);
    } else {
        maxpool_bwd_kernel<1><<<grid_for(n_targets * d), 256, 0, st>>>(feat, ld_f, out, ld_o, ties, g_out, ld_g, offsets,
                                                                     perm, n_targets, K, g_feat, ld_gf, d);
    }
    PU_LAUNCH_CHECK();
    return PU_OK;
}

size_t pu_point2prod_workspace_bytes(int Z, int X, int Y) {
    if (Z <= 0 || X <= 0 || Y <= 0) return 0;
    return (size_t)Z * X * Y * sizeof(int32_t);
}

int pu_point2prod(const float *probs, const int32_t *xyz_origin, const int32_t *point_idx, int n, int C, int Z, int X,
                  int Y, float *volume, void *workspace, size_t workspace_bytes, pu_stream_t stream) {
    if (!probs || !xyz_origin || !volume || n < 0 || C < 1 || Z < 1 || X < 1 || Y < 1) return PU_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < pu_point2prod_workspace_bytes(Z, X, Y)) return PU_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nvox = (size_t)Z * X * Y;
    PU_CUDA_TRY(cudaMemsetAsync(workspace, 0xFF, nvox * sizeof(int32_t), st));  // owner = -1
    PU_CUDA_TRY(cudaMemsetAsync(volume, 0, nvox * C * sizeof(float), st));      // np.zeros(volume_shape)
    if (n == 0) return PU_OK;
    p2v_owner_kernel<<<ceil_div(n, 256), 256, 0, st>>>(xyz_origin, point_idx, n, Z, X, Y, (int32_t *)workspace);
    PU_LAUNCH_CHECK();
    p2v_write_kernel<<<ceil_div(n, 256), 256, 0, st>>>(probs, xyz_origin, point_idx, n, C, Z, X, Y,
                                                       (const int32_t *)workspace, volume);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_point2label(const float *probs, const int32_t *xyz_origin, const int32_t *point_idx, int n, int C, int Z, int X,
                   int Y, int remap_from, int remap_to, unsigned char *labels, void *workspace, size_t workspace_bytes,
                   pu_stream_t stream) {
    if (!probs || !xyz_origin || !labels || n < 0 || C < 1 || C > 255 || Z < 1 || X < 1 || Y < 1) return PU_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < pu_point2prod_workspace_bytes(Z, X, Y)) return PU_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nvox = (size_t)Z * X * Y;
    PU_CUDA_TRY(cudaMemsetAsync(workspace, 0xFF, nvox * sizeof(int32_t), st));  // owner = -1
    PU_CUDA_TRY(cudaMemsetAsync(labels, 0, nvox, st));                          // argmax of an all-zero voxel
    if (n == 0) return PU_OK;
    p2v_owner_kernel<<<ceil_div(n, 256), 256, 0, st>>>(xyz_origin, point_idx, n, Z, X, Y, (int32_t *)workspace);
    PU_LAUNCH_CHECK();
    p2v_label_kernel<<<ceil_div(n, 256), 256, 0, st>>>(probs, xyz_origin, point_idx, n, C, Z, X, Y,
                                                       (const int32_t *)workspace, remap_from, remap_to, labels);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_volume_argmax(const float *volume, long long nvox, int C, int remap_from, int remap_to, unsigned char *labels,
                     pu_stream_t stream) {
    if (!volume || !labels || nvox < 0 || C < 1 || C > 255) return PU_ERR_INVALID_ARG;
    if (nvox == 0) return PU_OK;
    volume_argmax_kernel<<<grid_for(nvox), 256, 0, (cudaStream_t)stream>>>(volume, nvox, C, remap_from, remap_to, labels);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

}  // extern "C"
