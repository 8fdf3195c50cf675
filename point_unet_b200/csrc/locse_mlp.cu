// locse_mlp.cu -- the position branch of building_block as RECOMPUTE kernels (PointSegment/RandLANet.py:323-326):
//
//     f_xyz = relative_pos_encoding(xyz, neigh_idx)                       [B,N,K,10]   (:337-343)
//     f_xyz = conv2d(f_xyz, d_out/2, 'mlp1') + BN(0.99, 1e-6) + LeakyReLU [B,N,K,h]    (helper_tf_util.py:115-170)
//
// The unfused path writes the 10-channel LocSE rows, reads them back for the 10 -> h product, writes the pre-normalisation
// tensor y, reads it for the batch norm, and in the backward reads y twice more (reduce, apply), writes dy and reads dy and
// the LocSE rows again for the weight gradient: 32h + 120 bytes per (n,k) row that are pure round trips.  Here neither the
// LocSE rows nor y nor dy ever exist in memory:
//
//   * the batch statistics of y = xW + b follow from the 10 x 10 covariance of x:  mean_y = xbar W + b,
//     var_y[c] = W[:,c]^T Cov W[:,c]  (moment kernels: two passes over idx + L2-resident xyz, centred second moments);
//   * forward: one kernel recomputes x per row, forms  y - mean_y = (x - xbar) W  (the bias cancels, no y*scale - mean*scale
//     cancellation either) and stores lrelu(gamma*invstd*(y - mean_y) + beta) into the concat half (and the f_xyz copy);
//   * backward: one kernel reads the incoming gradient(s), recomputes x and y - mean_y, and reduces  sum g,  sum g*(y-mean_y)
//     and  G = (x - xbar)^T g  (g = gradient through the LeakyReLU).  Because xyz is data there is no dgrad, and the weight
//     gradient of the BN-wrapped product has a closed form in those sums:
//         dW[j,c] = gamma_c invstd_c ( G[j,c] - invstd_c^2 (sum_r g (y - mean_y))_c (Cov W)[j,c] ),   db = 0
//     (sum_r x_rj dy_rc with dy = gamma invstd (g - mean g - xhat mean(g xhat)); the mean-g term vanishes against the centring
//     of x, and sum_r x_rj (y - mean_y)_rc = M (Cov W)[j,c]).
//
// All reductions run in a fixed order (per-thread rows ascending, fixed shuffle / shared-memory trees, partials reduced in
// double in index order): results are bit-deterministic.
#include <type_traits>

#include "common.cuh"

namespace pu {
namespace locse {

constexpr int TILE = 256;       // (n,k) rows of one cloud staged per CTA iteration
constexpr int BWD_CTAS = 296;   // CTAs (partials) of the backward kernel: two per SM, all resident
constexpr int MOM_CTAS = 592;   // CTAs (partials) of the moment kernels
constexpr int NACC = 12;        // per-channel sums of the backward: sum g, sum g*yhat, G[0..9]

// The kernels read the cloud as PADDED points, xyz4 [B,N] float4 = (x, y, z, 0) (pack_xyz4_kernel): one 16-byte access per
// end point.  With the packed [N,3] layout every neighbour cost three scalar gathers, and the L1 data stage was the
// busiest unit of the moment and narrow-layer kernels (ncu: l1tex 70 %, 14.6 sectors per load request).
__global__ void __launch_bounds__(256) pack_xyz4_kernel(const float *__restrict__ xyz, long long n, float4 *__restrict__ xyz4) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) xyz4[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f);
}
// the two end points of row r: p = xyz4[n(r)], q = xyz4[j]
struct RowPts { float4 p, q; };
__device__ __forceinline__ RowPts load_pts(const float4 *__restrict__ xb, unsigned r, unsigned j, int K, int shiftK) {
    const unsigned n = shiftK >= 0 ? (r >> shiftK) : (r / (unsigned)K);
    RowPts c;
    c.p = xb[n];
    c.q = xb[j];
    return c;
}
// LocSE channels [ |p-q|, p-q, p, q ]  (same arithmetic as lfa::locse_kernel)
// The pad lanes (w = 0) are folded into the distance (+ 0*0, exact): a lane the arithmetic never reads is a dead register
// the moment the 128-bit load is ISSUED, ptxas hands it out as a temporary, and the first write to it waits for the load
// (write-after-write on an in-flight destination): ncu had 65 % of the stall samples of the narrow forward on one MUFU.RSQ
// whose destination was the .w register of the prefetch issued just above it.
__device__ __forceinline__ void locse_from_pts(const RowPts &c, float (&x)[10]) {
    const float rx = c.p.x - c.q.x, ry = c.p.y - c.q.y, rz = c.p.z - c.q.z, rw = c.p.w - c.q.w;
    x[0] = sqrtf(fmaf(rw, rw, rx * rx + ry * ry + rz * rz));
    x[1] = rx; x[2] = ry; x[3] = rz;
    x[4] = c.p.x; x[5] = c.p.y; x[6] = c.p.z;
    x[7] = c.q.x; x[8] = c.q.y; x[9] = c.q.z;
}

__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;   // src-size 0: nothing is read, the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void cp_async4_zfill(void *smem_dst, const void *gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

// A thread's rows are pipelined through SHARED MEMORY: the neighbour index of row i + 2D and the two end points of row
// i + D are requested with cp.async while row i is worked on (index -> gather is a dependent chain of a DRAM and an L2
// access, > 1000 cycles under load).  Each thread copies only what it will read itself, so the ring needs no barrier, and
// nothing in flight occupies a register.  The register pipelines tried first (index four rows / coordinates two rows
// ahead) ran into ptxas: rotating by moves reads a register that a load is still filling, an unused .w lane of a 128-bit
// load is reused as a temporary while the load is in flight, the sign extension of an index is placed right behind its
// load -- every time the kernel sat on the first use of a prefetched value (ncu: 7 long-scoreboard cycles per issue).
constexpr int PD = 4;   // rows between the request of a row's end points and their use
constexpr size_t ROWS_SMEM = (size_t)2 * PD * 256 * sizeof(unsigned) + (size_t)2 * PD * 256 * sizeof(float4);   // 40 KB
struct AsyncRows {
    unsigned *s_j;     // [2 PD][256]
    float4 *s_p, *s_q; // [PD][256] each
    const float4 *xb;
    const int32_t *ib;
    unsigned rpc, step, r;   // r: the row being worked on
    int K, shiftK, sp, sj;   // ring positions of row r: sp = i % PD, sj = i % (2 PD)

    __device__ __forceinline__ void issue_idx(unsigned row, int slot) {
        const bool valid = row < rpc;
        cp_async4_zfill(s_j + slot * 256 + threadIdx.x, valid ? ib + row : ib, valid);
    }
    __device__ __forceinline__ void issue_pts(unsigned row, int slot_j, int slot_p) {
        if (row < rpc) {
            const unsigned j = s_j[slot_j * 256 + threadIdx.x];
            const unsigned n = shiftK >= 0 ? (row >> shiftK) : (row / (unsigned)K);
            cp_async16_zfill(s_p + slot_p * 256 + threadIdx.x, xb + n, true);
            cp_async16_zfill(s_q + slot_p * 256 + threadIdx.x, xb + j, true);
        }
    }
    // carve the rings out of `smem` (ROWS_SMEM bytes, 16-byte aligned) and fill the pipeline.  `extra(row, slot)` may add
    // copies of its own for that row to the group (slot = position in a PD-deep ring): they land together with the end points.
    template <class F>
    __device__ __forceinline__ void start(void *smem, unsigned first, F extra) {
        s_p = reinterpret_cast<float4 *>(smem);
        s_q = s_p + PD * 256;
        s_j = reinterpret_cast<unsigned *>(s_q + PD * 256);
        r = first; sp = 0; sj = 0;
#pragma unroll
        for (int k = 0; k < PD; ++k) issue_idx(first + k * step, k);
        cp_async_commit();
        cp_async_wait<0>();
#pragma unroll
        for (int k = 0; k < PD; ++k) {   // group k = { end points of row k, index of row k + PD, extra(row k) }
            issue_pts(first + k * step, k, k);
            issue_idx(first + (k + PD) * step, k + PD);
            extra(first + k * step, k);
            cp_async_commit();
        }
    }
    // end points of the current row (valid once the oldest group has landed); its ring position is `sp`
    __device__ __forceinline__ RowPts current() {
        cp_async_wait<PD - 1>();
        RowPts c;
        c.p = s_p[sp * 256 + threadIdx.x];
        c.q = s_q[sp * 256 + threadIdx.x];
        return c;
    }
    // after the values of the current row have been used: refill the slots just read and step to the next row
    template <class F>
    __device__ __forceinline__ void advance(F extra) {
        issue_pts(r + PD * step, (sj + PD) & (2 * PD - 1), sp);      // its index arrived with the group just waited for
        issue_idx(r + 2 * PD * step, sj);
        extra(r + PD * step, sp);
        cp_async_commit();
        sp = (sp + 1) & (PD - 1);
        sj = (sj + 1) & (2 * PD - 1);
        r += step;
    }
};
struct NoExtra { __device__ __forceinline__ void operator()(unsigned, int) const {} };

// MODE 0: per-CTA sums of the 10 channels.  MODE 1: per-CTA sums of the centred products (x_i - xbar_i)(x_j - xbar_j),
// i <= j, 55 values in row-major upper-triangle order; xbar_j = sums[j] * inv_count, evaluated identically everywhere.
// A thread walks its rows with the neighbour index fetched four rows ahead and the coordinates two rows ahead (index ->
// gather is a dependent chain of a DRAM and an L2 access; with 55 accumulators only two CTAs fit an SM).
template <int MODE>
__global__ void __launch_bounds__(256) locse_moment_kernel(const float4 *__restrict__ xyz, const int32_t *__restrict__ idx, int N,
                                                           int K, int shiftK, unsigned rpc, const float *__restrict__ sums,
                                                           float inv_count, float *__restrict__ part) {
    constexpr int E = MODE == 0 ? 10 : 55;
    extern __shared__ __align__(16) unsigned char s_rows[];   // ROWS_SMEM
    __shared__ float s_red[8][E];
    float xbar[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) xbar[j] = MODE == 1 ? sums[j] * inv_count : 0.f;
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    const float4 *xb = xyz + (size_t)blockIdx.y * N;
    const int32_t *ib = idx + (size_t)blockIdx.y * rpc;
    AsyncRows rp;
    rp.xb = xb; rp.ib = ib; rp.rpc = rpc; rp.step = gridDim.x * 256u; rp.K = K; rp.shiftK = shiftK;
    rp.start(s_rows, blockIdx.x * 256u + threadIdx.x, NoExtra{});
    for (; rp.r < rpc; rp.advance(NoExtra{})) {
        const unsigned ru = rp.r;
        const RowPts pts = rp.current();
        float x[10];
        locse_from_pts(pts, x);
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 10; ++j) acc[j] += x[j];
        } else {
#pragma unroll
            for (int j = 0; j < 10; ++j) x[j] -= xbar[j];
            int e = 0;
#pragma unroll
            for (int i = 0; i < 10; ++i)
#pragma unroll
                for (int j = i; j < 10; ++j, ++e) acc[e] = fmaf(x[i], x[j], acc[e]);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        float v = acc[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][e] = v;
    }
    __syncthreads();
    if (threadIdx.x < E) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_red[w][threadIdx.x];
        part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * E + threadIdx.x] = s;
    }
}

// coef layout (floats): [0,h) scale | [h,2h) t | [2h,3h) invstd | [3h,4h) mean_y | [4h,5h) var_y | [5h,5h+10) xbar |
//                       [5h+10, 5h+110) Cov (10x10, biased)
// training: scale = gamma*invstd, t = beta, xbar / Cov from the moments; moving statistics updated when given.
// inference: scale = gamma / sqrt(moving_var + eps), t = beta + scale * (bias - moving_mean), xbar = 0.
__global__ void __launch_bounds__(256) locse_bn_prepare_kernel(const float *__restrict__ mom, float inv_count, const float *__restrict__ w,
                                                               int h, const float *__restrict__ bias, const float *__restrict__ gamma,
                                                               const float *__restrict__ beta, float eps, int training,
                                                               float *__restrict__ moving_mean, float *__restrict__ moving_var,
                                                               float momentum, float unbias, float *__restrict__ coef) {
    __shared__ double cov[10][10];
    __shared__ float xbar[10];
    if (threadIdx.x < 10) xbar[threadIdx.x] = training ? mom[threadIdx.x] * inv_count : 0.f;
    if (threadIdx.x < 55) {
        int i = 0, e = threadIdx.x;
        while (e >= 10 - i) { e -= 10 - i; ++i; }
        const int j = i + e;
        const double v = training ? (double)mom[10 + threadIdx.x] * (double)inv_count : 0.0;
        cov[i][j] = v;
        cov[j][i] = v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < h; c += 256) {
        float sc, t, is = 0.f, mf = 0.f, vf = 0.f;
        if (training) {
            double wc[10], mean = (double)bias[c], var = 0.0;
#pragma unroll
            for (int j = 0; j < 10; ++j) { wc[j] = (double)w[(size_t)j * h + c]; mean += (double)xbar[j] * wc[j]; }
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                double row = 0.0;
#pragma unroll
                for (int j = 0; j < 10; ++j) row += cov[i][j] * wc[j];
                var += wc[i] * row;
            }
            mf = (float)mean;
            vf = (float)(var > 0.0 ? var : 0.0);
            is = rsqrtf(vf + eps);
            sc = gamma[c] * is;
            t = beta[c];
            if (moving_mean) {
                moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * mf;
                moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * vf * unbias;
            }
        } else {
            mf = moving_mean[c];
            vf = moving_var[c];
            is = rsqrtf(vf + eps);
            sc = gamma[c] * is;
            t = fmaf(sc, bias[c] - mf, beta[c]);
        }
        coef[c] = sc;
        coef[h + c] = t;
        coef[2 * h + c] = is;
        coef[3 * h + c] = mf;
        coef[4 * h + c] = vf;
    }
    if (threadIdx.x < 10) coef[5 * h + threadIdx.x] = xbar[threadIdx.x];
    if (threadIdx.x < 100) coef[5 * h + 10 + threadIdx.x] = (float)cov[threadIdx.x / 10][threadIdx.x % 10];
}

// The centred LocSE rows of a tile are staged in shared memory, s_x[buffer][row][10], by a pipeline that hides the
// index -> coordinates chain (a DRAM access followed by an L2 gather) under the arithmetic of the tile before: while tile t
// is consumed from one buffer, every thread already holds the coordinates of its row of tile t+1 (requested before the
// arithmetic, stored into the other buffer after it) and the neighbour index of its row of tile t+2.  One barrier per tile.
struct Stager {
    const float4 *xb;
    const int32_t *ib;
    unsigned rpc, stride;
    int K, shiftK;
    unsigned j_next;  // neighbour index of my row of the NEXT tile
    RowPts pts;       // coordinates of my row of the next tile (valid between fetch() and store())
    bool valid;

    __device__ __forceinline__ void store_row(float *s_buf, const RowPts &c, const float *xbar) const {
        float x[10];
        locse_from_pts(c, x);
        float2 *o = reinterpret_cast<float2 *>(s_buf + threadIdx.x * 10);
#pragma unroll
        for (int j = 0; j < 10; j += 2) o[j >> 1] = make_float2(x[j] - xbar[j], x[j + 1] - xbar[j + 1]);
    }
    // stage tile r0 synchronously into s_buf and fetch the index for tile r0 + stride
    __device__ __forceinline__ void prologue(float *s_buf, unsigned r0, const float *xbar) {
        const unsigned r = r0 + threadIdx.x;
        if (r < rpc) store_row(s_buf, load_pts(xb, r, (unsigned)ib[r], K, shiftK), xbar);
        const unsigned rn = r + stride;
        j_next = rn < rpc ? (unsigned)ib[rn] : 0u;
    }
    // before the arithmetic of tile r0: request the coordinates of tile r0 + stride and the index of tile r0 + 2 stride
    __device__ __forceinline__ unsigned fetch(unsigned r0) {
        const unsigned rn = r0 + stride + threadIdx.x;
        valid = rn < rpc;
        if (valid) pts = load_pts(xb, rn, j_next, K, shiftK);
        const unsigned rnn = rn + stride;
        return rnn < rpc ? (unsigned)ib[rnn] : 0u;
    }
    // after the arithmetic: finish the next tile's rows into the other buffer
    __device__ __forceinline__ void store(float *s_buf, unsigned j_after, const float *xbar) {
        if (valid) store_row(s_buf, pts, xbar);
        j_next = j_after;
    }
};

// out[row, c] = lrelu( scale[c] * sum_j (x[row,j] - xbar[j]) W[j,c] + t[c] ), also into out2 when given.
// A thread owns ONE float4 column group (its 40 weights and 8 coefficients live in registers) and walks the rows of the
// staged tile; the cq threads that share a row read its 10 values as broadcasts.
__global__ void __launch_bounds__(256, 3) locse_mlp_fwd_kernel(const float4 *__restrict__ xyz, const int32_t *__restrict__ idx, int N,
                                                               int K, int shiftK, unsigned rpc, const float *__restrict__ w, int h,
                                                               const float *__restrict__ coef, float slope, float *__restrict__ out,
                                                               int ldo, float *__restrict__ out2, int ldo2) {
    __shared__ __align__(16) float s_x[2][TILE * 10];
    const int cq = h >> 2, rpb = 256 / cq;
    const int c = (threadIdx.x % cq) * 4, rl = threadIdx.x / cq;
    float wr[10][4];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const float4 v = *reinterpret_cast<const float4 *>(w + (size_t)j * h + c);
        wr[j][0] = v.x; wr[j][1] = v.y; wr[j][2] = v.z; wr[j][3] = v.w;
    }
    const float4 sc = *reinterpret_cast<const float4 *>(coef + c), tt = *reinterpret_cast<const float4 *>(coef + h + c);
    __shared__ float xbar[10];   // read at staging time only: kept out of the registers of the row loop
    if (threadIdx.x < 10) xbar[threadIdx.x] = coef[5 * h + threadIdx.x];
    __syncthreads();
    const size_t cloud_row0 = (size_t)blockIdx.y * rpc;
    Stager sg;
    sg.xb = xyz + (size_t)blockIdx.y * N;
    sg.ib = idx + cloud_row0;
    sg.rpc = rpc; sg.stride = gridDim.x * (unsigned)TILE; sg.K = K; sg.shiftK = shiftK;
    unsigned r0 = blockIdx.x * (unsigned)TILE;
    if (r0 >= rpc) return;
    sg.prologue(s_x[0], r0, xbar);
    __syncthreads();
    int cur = 0;
    for (; r0 < rpc; r0 += sg.stride) {
        const unsigned j_after = sg.fetch(r0);
        const float *sx = s_x[cur];
        const int nrows = (int)min((unsigned)TILE, rpc - r0);
        for (int rr = rl; rr < nrows; rr += rpb) {
            const float2 *xs = reinterpret_cast<const float2 *>(sx + rr * 10);
            float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 10; j += 2) {
                const float2 xv = xs[j >> 1];
#pragma unroll
                for (int k = 0; k < 4; ++k) a[k] = fmaf(xv.x, wr[j][k], a[k]);
#pragma unroll
                for (int k = 0; k < 4; ++k) a[k] = fmaf(xv.y, wr[j + 1][k], a[k]);
            }
            float4 z = make_float4(fmaf(a[0], sc.x, tt.x), fmaf(a[1], sc.y, tt.y), fmaf(a[2], sc.z, tt.z), fmaf(a[3], sc.w, tt.w));
            z.x = z.x > 0.f ? z.x : z.x * slope;
            z.y = z.y > 0.f ? z.y : z.y * slope;
            z.z = z.z > 0.f ? z.z : z.z * slope;
            z.w = z.w > 0.f ? z.w : z.w * slope;
            const size_t row = cloud_row0 + r0 + rr;
            *reinterpret_cast<float4 *>(out + row * ldo + c) = z;
            if (out2) *reinterpret_cast<float4 *>(out2 + row * ldo2 + c) = z;
        }
        sg.store(s_x[cur ^ 1], j_after, xbar);
        __syncthreads();
        cur ^= 1;
    }
}

// CTA reduction of the backward accumulators over the threads that share a column group (thread t owns group t % cq),
// fixed order: lanes (xor butterfly) -> warps (ascending); the CTA's partial goes to pb[a][c].  s_red: 8*32*16 floats.
__device__ __forceinline__ void reduce_acc_to_part(float (&acc)[NACC][4], int cq, float *s_red, float *pb, int h) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = cq; o < 32; o <<= 1) {
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[a][k] += __shfl_xor_sync(0xffffffffu, acc[a][k], o);
    }
    const int owners = cq < 32 ? cq : 32;          // lanes of a warp holding distinct column groups
    const int wpg = cq < 32 ? 1 : cq / 32;         // warps needed to cover all column groups once
#pragma unroll
    for (int round = 0; round < NACC / 4; ++round) {
        if (lane < owners) {
            float4 *o = reinterpret_cast<float4 *>(s_red + ((size_t)warp * 32 + lane) * 16);
#pragma unroll
            for (int a = 0; a < 4; ++a)
                o[a] = make_float4(acc[round * 4 + a][0], acc[round * 4 + a][1], acc[round * 4 + a][2], acc[round * 4 + a][3]);
        }
        __syncthreads();
        if (threadIdx.x < cq) {
            const int cg = threadIdx.x, w0 = cq < 32 ? 0 : cg / 32, ln = cq < 32 ? cg : cg % 32;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int wv = w0; wv < 8; wv += wpg) {
                    const float4 v = reinterpret_cast<const float4 *>(s_red + ((size_t)wv * 32 + ln) * 16)[a];
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
                *reinterpret_cast<float4 *>(pb + (size_t)(round * 4 + a) * h + cg * 4) = s;
            }
        }
        __syncthreads();
    }
}

constexpr int RING = 8;   // gradient rows in flight per thread
constexpr size_t BWD_SMEM = (size_t)RING * 2 * 256 * sizeof(float4);   // 64 KB; reused for the CTA reduction at the end
constexpr size_t BWD_DIRECT_SMEM = (size_t)PD * 2 * 256 * sizeof(float4) + ROWS_SMEM;   // 72 KB

// Backward sums.  part[cta][a][c], a = 0: sum g, 1: sum g*yhat, 2..11: sum (x_j - xbar_j) g, with yhat = (x - xbar) W and
// g = (dz [+ dz2]) * lrelu'(scale*yhat + t).
// The gradient rows arrive through a PER-THREAD cp.async ring, RING rows deep (a thread copies exactly the 16-byte chunks
// it will consume itself, so the ring needs no barrier): with 88 accumulator / weight registers only two CTAs fit an SM,
// and register prefetching one pair of rows ahead left the DRAM latency exposed (0.59 ms at level 0, 20 % of the copy
// peak; bytes in flight are what an HBM-bound kernel is made of).
__global__ void __launch_bounds__(256, 2) locse_mlp_bwd_kernel(const float4 *__restrict__ xyz, const int32_t *__restrict__ idx, int N,
                                                               int K, int shiftK, unsigned rpc, const float *__restrict__ w, int h,
                                                               const float *__restrict__ coef, float slope,
                                                               const float *__restrict__ dz, int ldz, const float *__restrict__ dz2,
                                                               int ldz2, float *__restrict__ part) {
    __shared__ __align__(16) float s_x[2][TILE * 10];
    extern __shared__ __align__(16) float4 s_ring[];   // [RING][2][256]
    float *s_red = reinterpret_cast<float *>(s_ring);  // [8 warps][32 lanes][16] after the row loop
    const int cq = h >> 2, rpb = 256 / cq;
    const int c = (threadIdx.x % cq) * 4, rl = threadIdx.x / cq;
    float wr[10][4];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const float4 v = *reinterpret_cast<const float4 *>(w + (size_t)j * h + c);
        wr[j][0] = v.x; wr[j][1] = v.y; wr[j][2] = v.z; wr[j][3] = v.w;
    }
    const float4 sc4 = *reinterpret_cast<const float4 *>(coef + c), tt4 = *reinterpret_cast<const float4 *>(coef + h + c);
    const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, tt[4] = {tt4.x, tt4.y, tt4.z, tt4.w};
    __shared__ float xbar[10];   // read at staging time only: kept out of the registers of the row loop
    if (threadIdx.x < 10) xbar[threadIdx.x] = coef[5 * h + threadIdx.x];
    __syncthreads();
    float acc[NACC][4];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[a][k] = 0.f;
    const size_t cloud_row0 = (size_t)blockIdx.y * rpc;
    const float *dzc = dz + cloud_row0 * ldz + c;
    const float *dz2c = dz2 ? dz2 + cloud_row0 * ldz2 + c : nullptr;

    const float *sx = nullptr;
    auto consume = [&](int rr, const float4 &g4) {
        const float2 *xs = reinterpret_cast<const float2 *>(sx + rr * 10);
        float x[10];
#pragma unroll
        for (int j = 0; j < 10; j += 2) {
            const float2 xv = xs[j >> 1];
            x[j] = xv.x; x[j + 1] = xv.y;
        }
        float yh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 10; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) yh[k] = fmaf(x[j], wr[j][k], yh[k]);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
        float gz[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float z = fmaf(yh[k], sc[k], tt[k]);
            gz[k] = z > 0.f ? g[k] : g[k] * slope;
            acc[0][k] += gz[k];
            acc[1][k] = fmaf(gz[k], yh[k], acc[1][k]);
        }
#pragma unroll
        for (int j = 0; j < 10; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[2 + j][k] = fmaf(x[j], gz[k], acc[2 + j][k]);
    };

    Stager sg;
    sg.xb = xyz + (size_t)blockIdx.y * N;
    sg.ib = idx + cloud_row0;
    sg.rpc = rpc; sg.stride = gridDim.x * (unsigned)TILE; sg.K = K; sg.shiftK = shiftK;
    unsigned r0 = blockIdx.x * (unsigned)TILE;

    // ring entry e of this thread = its e-th row in consumption order: tiles r0, r0 + stride, ...; rows rl, rl + rpb, ...
    // of each tile (always TILE / rpb entries per tile; rows beyond the cloud are zero-filled and skipped)
    unsigned pr0 = r0;
    int prr = rl;
    auto issue = [&](int slot) {
        const unsigned r = pr0 + (unsigned)prr;
        const bool valid = pr0 < rpc && r < rpc;
        float4 *dst = s_ring + (size_t)slot * 512 + threadIdx.x;
        cp_async16_zfill(dst, valid ? dzc + (size_t)r * ldz : dz, valid);
        if (dz2) cp_async16_zfill(dst + 256, valid ? dz2c + (size_t)r * ldz2 : dz2, valid);
        cp_async_commit();
        prr += rpb;
        if (prr >= TILE) { prr = rl; pr0 += sg.stride; }
    };
#pragma unroll
    for (int e = 0; e < RING; ++e) issue(e);
    if (r0 < rpc) sg.prologue(s_x[0], r0, xbar);
    __syncthreads();
    int cur = 0, slot = 0;
    for (; r0 < rpc; r0 += sg.stride) {
        const unsigned j_after = sg.fetch(r0);
        sx = s_x[cur];
        for (int rr = rl; rr < TILE; rr += rpb) {
            cp_async_wait<RING - 1>();
            float4 g = s_ring[(size_t)slot * 512 + threadIdx.x];
            if (dz2) {
                const float4 g2 = s_ring[(size_t)slot * 512 + 256 + threadIdx.x];
                g.x += g2.x; g.y += g2.y; g.z += g2.z; g.w += g2.w;
            }
            if (r0 + (unsigned)rr < rpc) consume(rr, g);
            issue(slot);   // refill the slot just read (program order after the arithmetic that used it)
            slot = (slot + 1) & (RING - 1);
        }
        sg.store(s_x[cur ^ 1], j_after, xbar);
        __syncthreads();
        cur ^= 1;
    }
    cp_async_wait<0>();
    __syncthreads();

    reduce_acc_to_part(acc, cq, s_red, part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * NACC * h, h);
}

// ---- narrow layers (h <= 16: level 0) --------------------------------------------------------------------------------
// With CQ = h/4 <= 4 threads per row a staged tile gives a thread only CQ rows of arithmetic per barrier and per
// index -> coordinates round trip (ncu: 37 % issue utilisation, long-scoreboard + barrier stalls).  Here every thread
// recomputes the LocSE channels of its own row (CQ-fold redundant, ~35 instructions) and nothing is shared: no staging, no
// barrier, and the loads of a thread are pipelined across ITS rows (index four rows ahead, coordinates two rows ahead).
// forward of the narrow layers: ONE thread per row, all h = 4 CQ channels (the pipeline bookkeeping of a row -- ~110 of the
// ~210 instructions a (row, column group) cost with CQ threads per row -- is paid once per row); the weights are read from
// shared memory, every lane the same address (broadcast), the two or four float4 stores of a row are contiguous.
template <int CQ>
__global__ void __launch_bounds__(256, 3) locse_mlp_fwd_direct_kernel(const float4 *__restrict__ xyz, const int32_t *__restrict__ idx,
                                                                      int N, int K, int shiftK, unsigned rpc,
                                                                      const float *__restrict__ w, const float *__restrict__ coef,
                                                                      float slope, float *__restrict__ out, int ldo,
                                                                      float *__restrict__ out2, int ldo2) {
    constexpr int h = CQ * 4;
    extern __shared__ __align__(16) unsigned char s_rows[];   // ROWS_SMEM
    __shared__ __align__(16) float s_w[10 * h];
    __shared__ __align__(16) float s_c[2 * h];                // scale | t
    __shared__ float s_xbar[10];
    for (int i = threadIdx.x; i < 10 * h; i += 256) s_w[i] = w[i];
    for (int i = threadIdx.x; i < 2 * h; i += 256) s_c[i] = coef[i];
    if (threadIdx.x < 10) s_xbar[threadIdx.x] = coef[5 * h + threadIdx.x];
    __syncthreads();
    const size_t cloud_row0 = (size_t)blockIdx.y * rpc;
    AsyncRows rp;
    rp.xb = xyz + (size_t)blockIdx.y * N;
    rp.ib = idx + cloud_row0;
    rp.rpc = rpc; rp.step = gridDim.x * 256u; rp.K = K; rp.shiftK = shiftK;
    rp.start(s_rows, blockIdx.x * 256u + threadIdx.x, NoExtra{});
    for (; rp.r < rpc; rp.advance(NoExtra{})) {
        const unsigned ru = rp.r;
        const RowPts pts = rp.current();
        float x[10];
        locse_from_pts(pts, x);
        float a[CQ][4];
#pragma unroll
        for (int q = 0; q < CQ; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k) a[q][k] = 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const float xc = x[j] - s_xbar[j];
#pragma unroll
            for (int q = 0; q < CQ; ++q) {
                const float4 wv = *reinterpret_cast<const float4 *>(s_w + j * h + q * 4);
                a[q][0] = fmaf(xc, wv.x, a[q][0]); a[q][1] = fmaf(xc, wv.y, a[q][1]);
                a[q][2] = fmaf(xc, wv.z, a[q][2]); a[q][3] = fmaf(xc, wv.w, a[q][3]);
            }
        }
        const size_t row = cloud_row0 + ru;
#pragma unroll
        for (int q = 0; q < CQ; ++q) {
            const float4 sc = *reinterpret_cast<const float4 *>(s_c + q * 4), tt = *reinterpret_cast<const float4 *>(s_c + h + q * 4);
            float4 z = make_float4(fmaf(a[q][0], sc.x, tt.x), fmaf(a[q][1], sc.y, tt.y), fmaf(a[q][2], sc.z, tt.z),
                                   fmaf(a[q][3], sc.w, tt.w));
            z.x = z.x > 0.f ? z.x : z.x * slope;
            z.y = z.y > 0.f ? z.y : z.y * slope;
            z.z = z.z > 0.f ? z.z : z.z * slope;
            z.w = z.w > 0.f ? z.w : z.w * slope;
            *reinterpret_cast<float4 *>(out + row * ldo + q * 4) = z;
            if (out2) *reinterpret_cast<float4 *>(out2 + row * ldo2 + q * 4) = z;
        }
    }
}

// backward sums of the narrow layers: rows as above, gradients through the per-thread cp.async ring; the 40 weights of a
// column group are read from shared memory (CQ distinct addresses per warp: broadcasts) to leave the registers to the
// 48 accumulators and the coordinate pipeline.
template <int CQ>
__global__ void __launch_bounds__(256, 2) locse_mlp_bwd_direct_kernel(const float4 *__restrict__ xyz, const int32_t *__restrict__ idx,
                                                                      int N, int K, int shiftK, unsigned rpc,
                                                                      const float *__restrict__ w, const float *__restrict__ coef,
                                                                      float slope, const float *__restrict__ dz, int ldz,
                                                                      const float *__restrict__ dz2, int ldz2,
                                                                      float *__restrict__ part) {
    constexpr int h = CQ * 4, RPB = 256 / CQ;
    extern __shared__ __align__(16) float4 s_ring[];   // gradients [PD][2][256] float4, then the row rings (ROWS_SMEM)
    float *s_red = reinterpret_cast<float *>(s_ring);
    __shared__ __align__(16) float s_w[10 * h];
    __shared__ float s_xbar[10];
    for (int i = threadIdx.x; i < 10 * h; i += 256) s_w[i] = w[i];
    if (threadIdx.x < 10) s_xbar[threadIdx.x] = coef[5 * h + threadIdx.x];
    __syncthreads();
    const int c = (threadIdx.x % CQ) * 4;
    const float4 sc4 = *reinterpret_cast<const float4 *>(coef + c), tt4 = *reinterpret_cast<const float4 *>(coef + h + c);
    const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, tt[4] = {tt4.x, tt4.y, tt4.z, tt4.w};
    float acc[NACC][4];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[a][k] = 0.f;
    const size_t cloud_row0 = (size_t)blockIdx.y * rpc;
    const float *dzc = dz + cloud_row0 * ldz + c;
    const float *dz2c = dz2 ? dz2 + cloud_row0 * ldz2 + c : nullptr;
    AsyncRows rp;
    rp.xb = xyz + (size_t)blockIdx.y * N;
    rp.ib = idx + cloud_row0;
    rp.rpc = rpc; rp.step = gridDim.x * (unsigned)RPB; rp.K = K; rp.shiftK = shiftK;
    const unsigned first = blockIdx.x * (unsigned)RPB + threadIdx.x / CQ;
    // the gradient row(s) of a row travel in the same cp.async group as its end points: s_ring [PD][2][256] float4
    auto grad = [&](unsigned row, int slot) {
        const bool valid = row < rpc;
        float4 *dst = s_ring + (size_t)slot * 512 + threadIdx.x;
        cp_async16_zfill(dst, valid ? dzc + (size_t)row * ldz : dz, valid);
        if (dz2) cp_async16_zfill(dst + 256, valid ? dz2c + (size_t)row * ldz2 : dz2, valid);
    };
    rp.start(s_ring + PD * 2 * 256, first, grad);
    for (; rp.r < rpc; rp.advance(grad)) {
        const RowPts pts = rp.current();
        float4 g4 = s_ring[(size_t)rp.sp * 512 + threadIdx.x];
        if (dz2) {
            const float4 g2 = s_ring[(size_t)rp.sp * 512 + 256 + threadIdx.x];
            g4.x += g2.x; g4.y += g2.y; g4.z += g2.z; g4.w += g2.w;
        }
        float x[10];
        locse_from_pts(pts, x);
        float yh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            x[j] -= s_xbar[j];
            const float4 wv = *reinterpret_cast<const float4 *>(s_w + j * h + c);
            yh[0] = fmaf(x[j], wv.x, yh[0]); yh[1] = fmaf(x[j], wv.y, yh[1]);
            yh[2] = fmaf(x[j], wv.z, yh[2]); yh[3] = fmaf(x[j], wv.w, yh[3]);
        }
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
        float gz[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float z = fmaf(yh[k], sc[k], tt[k]);
            gz[k] = z > 0.f ? g[k] : g[k] * slope;
            acc[0][k] += gz[k];
            acc[1][k] = fmaf(gz[k], yh[k], acc[1][k]);
        }
#pragma unroll
        for (int j = 0; j < 10; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[2 + j][k] = fmaf(x[j], gz[k], acc[2 + j][k]);
    }
    cp_async_wait<0>();
    __syncthreads();
    reduce_acc_to_part(acc, CQ, s_red, part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * NACC * h, h);
}

// One CTA per channel: reduce the partials in double (fixed order), then the closed forms of the header comment.
__global__ void __launch_bounds__(256) locse_mlp_bwd_finalize_kernel(const float *__restrict__ part, int chunks, int h,
                                                                     const float *__restrict__ w, const float *__restrict__ coef,
                                                                     const float *__restrict__ gamma,
                                                                     const float *__restrict__ bias, int training,
                                                                     float *__restrict__ dw, int accumulate_dw,
                                                                     float *__restrict__ dbias, float *__restrict__ dgamma,
                                                                     float *__restrict__ dbeta) {
    constexpr int G = 21;  // chunk groups: 21 * 12 = 252 threads
    __shared__ double red[G][NACC];
    __shared__ double tot[NACC];
    const int c = blockIdx.x;
    const int a = threadIdx.x % NACC, g = threadIdx.x / NACC;
    if (g < G) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int ch = g;
        for (; ch + 3 * G < chunks; ch += 4 * G) {
            const float v0 = part[((size_t)ch * NACC + a) * h + c], v1 = part[((size_t)(ch + G) * NACC + a) * h + c],
                        v2 = part[((size_t)(ch + 2 * G) * NACC + a) * h + c], v3 = part[((size_t)(ch + 3 * G) * NACC + a) * h + c];
            s0 += (double)v0; s1 += (double)v1; s2 += (double)v2; s3 += (double)v3;
        }
        for (; ch < chunks; ch += G) s0 += (double)part[((size_t)ch * NACC + a) * h + c];
        red[g][a] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        double s = 0.0;
        for (int gg = 0; gg < G; ++gg) s += red[gg][threadIdx.x];
        tot[threadIdx.x] = s;
    }
    __syncthreads();
    const double is = (double)coef[2 * h + c], ga = (double)gamma[c];
    const double s = tot[0], q = tot[1];
    if (threadIdx.x == 0) {
        // inference-mode statistics (moving mean in coef[3h..]): yhat was x W, so y - mean = yhat + bias - moving_mean
        const double shift = training ? 0.0 : (double)bias[c] - (double)coef[3 * h + c];
        dgamma[c] = (float)((q + shift * s) * is);
        dbeta[c] = (float)s;
        // a bias in front of a training-mode batch norm has an identically zero gradient
        if (dbias) dbias[c] = training ? 0.f : (float)(ga * is * s);
    }
    if (threadIdx.x < 10) {
        const int j = threadIdx.x;
        double v = tot[2 + j];
        if (training) {
            double cw = 0.0;
#pragma unroll
            for (int i = 0; i < 10; ++i) cw += (double)coef[5 * h + 10 + j * 10 + i] * (double)w[(size_t)i * h + c];
            v -= is * is * q * cw;   // (invstd^2 q / M) * sum_r x_rj (y - mean_y)_rc,  the sum being M (Cov W)[j,c]
        }
        const float r = (float)(ga * is * v);
        float *o = dw + (size_t)j * h + c;
        *o = accumulate_dw ? *o + r : r;
    }
}

}  // namespace locse
}  // namespace pu

using namespace pu;
using namespace pu::locse;

namespace {
struct Geom {
    int shiftK;
    unsigned rpc;
};
int geom(int B, int N, int K, Geom *g) {
    if (B < 0 || N < 0 || K < 1) return PU_ERR_INVALID_ARG;
    const long long rpc = (long long)N * K;
    if (rpc >= (1ll << 31) - 256 * 4096ll || B > 65535) return PU_ERR_UNSUPPORTED;
    g->rpc = (unsigned)rpc;
    g->shiftK = -1;
    if ((K & (K - 1)) == 0) { g->shiftK = 0; while ((1 << g->shiftK) < K) ++g->shiftK; }
    return PU_OK;
}
inline unsigned grid_x(unsigned rpc, int B, int budget) {
    const unsigned tiles = (unsigned)((rpc + TILE - 1) / TILE);
    unsigned gx = (unsigned)(budget / (B > 0 ? B : 1));
    if (gx < 1) gx = 1;
    return tiles < gx ? tiles : gx;
}
inline bool width_ok(int h) { return h >= 4 && h <= 1024 && (h & (h - 1)) == 0; }
}  // namespace

extern "C" {

int pu_locse_mlp_supported(int K, int h) { return (K >= 1 && width_ok(h)) ? 1 : 0; }

size_t pu_locse_mlp_workspace_bytes(int h) {
    const size_t mom = (size_t)MOM_CTAS * 55 * sizeof(float);
    const size_t bwd = (size_t)BWD_CTAS * NACC * (size_t)(h > 0 ? h : 1) * sizeof(float);
    return align_up(mom > bwd ? mom : bwd, 256);
}

int pu_locse_pack_xyz(const float *xyz, long long n_points, float *xyz4, pu_stream_t stream) {
    if (!xyz || !xyz4 || n_points < 0 || (((uintptr_t)xyz4) & 15)) return PU_ERR_INVALID_ARG;
    if (n_points == 0) return PU_OK;
    pack_xyz4_kernel<<<ceil_div(n_points, 256), 256, 0, (cudaStream_t)stream>>>(xyz, n_points, reinterpret_cast<float4 *>(xyz4));
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_locse_moments(const float *xyz4, const int32_t *idx, int B, int N, int K, float *mom, void *workspace,
                     size_t workspace_bytes, pu_stream_t stream) {
    if (!xyz4 || !idx || !mom || (((uintptr_t)xyz4) & 15)) return PU_ERR_INVALID_ARG;
    const float4 *xyz = reinterpret_cast<const float4 *>(xyz4);
    Geom g;
    const int rc = geom(B, N, K, &g);
    if (rc != PU_OK) return rc;
    if (B == 0 || g.rpc == 0) return PU_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < pu_locse_mlp_workspace_bytes(1)) return PU_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float *part = reinterpret_cast<float *>(workspace);
    const dim3 grid(grid_x(g.rpc, B, MOM_CTAS), (unsigned)B);
    const int chunks = (int)(grid.x * grid.y);
    const float inv_count = (float)(1.0 / ((double)B * (double)g.rpc));
    locse_moment_kernel<0><<<grid, 256, ROWS_SMEM, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, nullptr, 0.f, part);
    PU_LAUNCH_CHECK();
    launch_reduce_parts(part, chunks, 10, mom, 0, st);
    PU_LAUNCH_CHECK();
    locse_moment_kernel<1><<<grid, 256, ROWS_SMEM, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, mom, inv_count, part);
    PU_LAUNCH_CHECK();
    launch_reduce_parts(part, chunks, 55, mom + 10, 0, st);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_locse_bn_prepare(const float *mom, long long count, const float *w, int h, const float *bias, const float *gamma,
                        const float *beta, float eps, int training, float *moving_mean, float *moving_var, float momentum,
                        float unbias, float *coef, pu_stream_t stream) {
    if (!w || !bias || !gamma || !beta || !coef || h < 1 || count < 1) return PU_ERR_INVALID_ARG;
    if (training && !mom) return PU_ERR_INVALID_ARG;
    if (!training && (!moving_mean || !moving_var)) return PU_ERR_INVALID_ARG;
    const float inv_count = (float)(1.0 / (double)count);
    locse_bn_prepare_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(mom, inv_count, w, h, bias, gamma, beta, eps, training ? 1 : 0,
                                                                 moving_mean, moving_var, momentum, unbias, coef);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_locse_mlp_fwd(const float *xyz4, const int32_t *idx, int B, int N, int K, const float *w, int h, const float *coef,
                     float slope, float *out, int ldo, float *out2, int ldo2, pu_stream_t stream) {
    if (!xyz4 || !idx || !w || !coef || !out || ldo < h || (out2 && ldo2 < h) || (((uintptr_t)xyz4) & 15)) return PU_ERR_INVALID_ARG;
    const float4 *xyz = reinterpret_cast<const float4 *>(xyz4);
    if (!width_ok(h)) return PU_ERR_UNSUPPORTED;
    if ((((uintptr_t)out) & 15) || (ldo & 3) || (out2 && ((((uintptr_t)out2) & 15) || (ldo2 & 3))) || (((uintptr_t)w) & 15) ||
        (((uintptr_t)coef) & 15))
        return PU_ERR_INVALID_ARG;
    Geom g;
    const int rc = geom(B, N, K, &g);
    if (rc != PU_OK) return rc;
    if (B == 0 || g.rpc == 0) return PU_OK;
    if (h <= 16) {   // narrow layers: one row per thread group, no staging
        const int rpb = 256;        // rows per CTA iteration: one thread per row
        unsigned gx = (unsigned)((g.rpc + rpb - 1) / rpb), cap = (unsigned)(kNumSMs * 3 * 4 / B);
        if (cap < 1) cap = 1;
        const dim3 grid(gx < cap ? gx : cap, (unsigned)B);
        cudaStream_t st = (cudaStream_t)stream;
        if (h == 4) locse_mlp_fwd_direct_kernel<1><<<grid, 256, ROWS_SMEM, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, w, coef, slope, out, ldo, out2, ldo2);
        else if (h == 8) locse_mlp_fwd_direct_kernel<2><<<grid, 256, ROWS_SMEM, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, w, coef, slope, out, ldo, out2, ldo2);
        else locse_mlp_fwd_direct_kernel<4><<<grid, 256, ROWS_SMEM, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, w, coef, slope, out, ldo, out2, ldo2);
        PU_LAUNCH_CHECK();
        return PU_OK;
    }
    const dim3 grid(grid_x(g.rpc, B, kNumSMs * 8), (unsigned)B);
    locse_mlp_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, idx, N, K, g.shiftK, g.rpc, w, h, coef, slope, out, ldo, out2,
                                                                 ldo2);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_locse_mlp_bwd(const float *xyz4, const int32_t *idx, int B, int N, int K, const float *w, int h, const float *coef,
                     const float *gamma, const float *bias, int training, float slope, const float *dz, int ldz, const float *dz2, int ldz2,
                     float *dw, int accumulate_dw, float *dbias, float *dgamma, float *dbeta, void *workspace,
                     size_t workspace_bytes, pu_stream_t stream) {
    if (!xyz4 || !idx || !w || !coef || !gamma || !bias || !dz || !dw || !dgamma || !dbeta || ldz < h || (dz2 && ldz2 < h) ||
        (((uintptr_t)xyz4) & 15))
        return PU_ERR_INVALID_ARG;
    const float4 *xyz = reinterpret_cast<const float4 *>(xyz4);
    if (!width_ok(h)) return PU_ERR_UNSUPPORTED;
    if ((((uintptr_t)dz) & 15) || (ldz & 3) || (dz2 && ((((uintptr_t)dz2) & 15) || (ldz2 & 3))) || (((uintptr_t)w) & 15) ||
        (((uintptr_t)coef) & 15))
        return PU_ERR_INVALID_ARG;
    Geom g;
    const int rc = geom(B, N, K, &g);
    if (rc != PU_OK) return rc;
    if (B == 0 || g.rpc == 0) return PU_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < pu_locse_mlp_workspace_bytes(h) || (((uintptr_t)workspace) & 15)) return PU_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float *part = reinterpret_cast<float *>(workspace);
    dim3 grid(grid_x(g.rpc, B, BWD_CTAS), (unsigned)B);
    static bool attr_set[64] = {};   // per device: the ring needs more than the default 48 KB of dynamic shared memory
    int dev = 0;
    PU_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        PU_CUDA_TRY(cudaFuncSetAttribute(locse_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
        PU_CUDA_TRY(cudaFuncSetAttribute(locse_mlp_bwd_direct_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_DIRECT_SMEM));
        PU_CUDA_TRY(cudaFuncSetAttribute(locse_mlp_bwd_direct_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_DIRECT_SMEM));
        PU_CUDA_TRY(cudaFuncSetAttribute(locse_mlp_bwd_direct_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_DIRECT_SMEM));
        attr_set[dev] = true;
    }
    if (h <= 16) {
        const int rpb = 1024 / h;
        unsigned gx = (unsigned)((g.rpc + rpb - 1) / rpb), cap = (unsigned)(BWD_CTAS / B);
        if (cap < 1) cap = 1;
        grid.x = gx < cap ? gx : cap;
        if (h == 4) locse_mlp_bwd_direct_kernel<1><<<grid, 256, BWD_DIRECT_SMEM, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, w, coef, slope, dz, ldz, dz2, ldz2, part);
        else if (h == 8) locse_mlp_bwd_direct_kernel<2><<<grid, 256, BWD_DIRECT_SMEM, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, w, coef, slope, dz, ldz, dz2, ldz2, part);
        else locse_mlp_bwd_direct_kernel<4><<<grid, 256, BWD_DIRECT_SMEM, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, w, coef, slope, dz, ldz, dz2, ldz2, part);
    } else {
        locse_mlp_bwd_kernel<<<grid, 256, BWD_SMEM, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, w, h, coef, slope, dz, ldz, dz2, ldz2, part);
    }
    PU_LAUNCH_CHECK();
    locse_mlp_bwd_finalize_kernel<<<h, 256, 0, st>>>(part, (int)(grid.x * grid.y), h, w, coef, gamma, bias, training ? 1 : 0, dw,
                                                     accumulate_dw, dbias, dgamma, dbeta);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

}  // extern "C"
