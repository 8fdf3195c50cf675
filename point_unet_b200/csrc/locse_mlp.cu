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
#include "common.cuh"

namespace pu {
namespace locse {

constexpr int TILE = 256;       // (n,k) rows of one cloud staged per CTA iteration
constexpr int BWD_CTAS = 296;   // CTAs (partials) of the backward kernel: two per SM, all resident
constexpr int MOM_CTAS = 592;   // CTAs (partials) of the moment kernels
constexpr int NACC = 12;        // per-channel sums of the backward: sum g, sum g*yhat, G[0..9]

// LocSE channels of row r of a cloud: [ |p-q|, p-q, p, q ]  (same arithmetic as lfa::locse_kernel)
__device__ __forceinline__ void locse_x(const float *__restrict__ xb, const int32_t *__restrict__ ib, unsigned r, int K,
                                        int shiftK, float (&x)[10]) {
    const unsigned n = shiftK >= 0 ? (r >> shiftK) : (r / (unsigned)K);
    const int j = ib[r];
    const float *p = xb + (size_t)n * 3;
    const float *q = xb + (size_t)j * 3;
    const float px = p[0], py = p[1], pz = p[2], qx = q[0], qy = q[1], qz = q[2];
    const float rx = px - qx, ry = py - qy, rz = pz - qz;
    x[0] = sqrtf(rx * rx + ry * ry + rz * rz);
    x[1] = rx; x[2] = ry; x[3] = rz;
    x[4] = px; x[5] = py; x[6] = pz;
    x[7] = qx; x[8] = qy; x[9] = qz;
}

// MODE 0: per-CTA sums of the 10 channels.  MODE 1: per-CTA sums of the centred products (x_i - xbar_i)(x_j - xbar_j),
// i <= j, 55 values in row-major upper-triangle order; xbar_j = sums[j] * inv_count, evaluated identically everywhere.
template <int MODE>
__global__ void __launch_bounds__(256) locse_moment_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ idx, int N,
                                                           int K, int shiftK, unsigned rpc, const float *__restrict__ sums,
                                                           float inv_count, float *__restrict__ part) {
    constexpr int E = MODE == 0 ? 10 : 55;
    __shared__ float s_red[8][E];
    float xbar[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) xbar[j] = MODE == 1 ? sums[j] * inv_count : 0.f;
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    const float *xb = xyz + (size_t)blockIdx.y * N * 3;
    const int32_t *ib = idx + (size_t)blockIdx.y * rpc;
    for (unsigned r = blockIdx.x * 256u + threadIdx.x; r < rpc; r += gridDim.x * 256u) {
        float x[10];
        locse_x(xb, ib, r, K, shiftK, x);
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

// stage the centred LocSE rows of one tile: s_x[row][10]
__device__ __forceinline__ void stage_tile(float *s_x, const float *__restrict__ xb, const int32_t *__restrict__ ib, unsigned r0,
                                           unsigned rpc, int K, int shiftK, const float (&xbar)[10]) {
    const unsigned r = r0 + threadIdx.x;
    if (r < rpc) {
        float x[10];
        locse_x(xb, ib, r, K, shiftK, x);
        float2 *o = reinterpret_cast<float2 *>(s_x + threadIdx.x * 10);
#pragma unroll
        for (int j = 0; j < 10; j += 2) o[j >> 1] = make_float2(x[j] - xbar[j], x[j + 1] - xbar[j + 1]);
    }
}

// out[row, c] = lrelu( scale[c] * sum_j (x[row,j] - xbar[j]) W[j,c] + t[c] ), also into out2 when given.
// A thread owns ONE float4 column group (its 40 weights and 8 coefficients live in registers) and walks the rows of the
// staged tile; the cq threads that share a row read its 10 values as broadcasts.
__global__ void __launch_bounds__(256) locse_mlp_fwd_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ idx, int N, int K,
                                                            int shiftK, unsigned rpc, const float *__restrict__ w, int h,
                                                            const float *__restrict__ coef, float slope, float *__restrict__ out,
                                                            int ldo, float *__restrict__ out2, int ldo2) {
    __shared__ __align__(16) float s_x[TILE * 10];
    const int cq = h >> 2, rpb = 256 / cq;
    const int c = (threadIdx.x % cq) * 4, rl = threadIdx.x / cq;
    float wr[10][4];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const float4 v = *reinterpret_cast<const float4 *>(w + (size_t)j * h + c);
        wr[j][0] = v.x; wr[j][1] = v.y; wr[j][2] = v.z; wr[j][3] = v.w;
    }
    const float4 sc = *reinterpret_cast<const float4 *>(coef + c), tt = *reinterpret_cast<const float4 *>(coef + h + c);
    float xbar[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) xbar[j] = coef[5 * h + j];
    const float *xb = xyz + (size_t)blockIdx.y * N * 3;
    const int32_t *ib = idx + (size_t)blockIdx.y * rpc;
    const size_t cloud_row0 = (size_t)blockIdx.y * rpc;
    for (unsigned r0 = blockIdx.x * (unsigned)TILE; r0 < rpc; r0 += gridDim.x * (unsigned)TILE) {
        stage_tile(s_x, xb, ib, r0, rpc, K, shiftK, xbar);
        __syncthreads();
        const int nrows = (int)min((unsigned)TILE, rpc - r0);
        for (int rr = rl; rr < nrows; rr += rpb) {
            const float2 *xs = reinterpret_cast<const float2 *>(s_x + rr * 10);
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
        __syncthreads();
    }
}

// Backward sums.  part[cta][a][c], a = 0: sum g, 1: sum g*yhat, 2..11: sum (x_j - xbar_j) g, with yhat = (x - xbar) W and
// g = (dz [+ dz2]) * lrelu'(scale*yhat + t).
__global__ void __launch_bounds__(256, 2) locse_mlp_bwd_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ idx, int N,
                                                               int K, int shiftK, unsigned rpc, const float *__restrict__ w, int h,
                                                               const float *__restrict__ coef, float slope,
                                                               const float *__restrict__ dz, int ldz, const float *__restrict__ dz2,
                                                               int ldz2, float *__restrict__ part) {
    __shared__ __align__(16) float s_x[TILE * 10];
    __shared__ __align__(16) float s_red[8 * 32 * 16];
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
    float xbar[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) xbar[j] = coef[5 * h + j];
    float acc[NACC][4];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[a][k] = 0.f;
    const float *xb = xyz + (size_t)blockIdx.y * N * 3;
    const int32_t *ib = idx + (size_t)blockIdx.y * rpc;
    const size_t cloud_row0 = (size_t)blockIdx.y * rpc;

    auto load_g = [&](size_t row) {
        float4 g = ld_stream_f4(reinterpret_cast<const float4 *>(dz + row * ldz + c));
        if (dz2) {
            const float4 g2 = ld_stream_f4(reinterpret_cast<const float4 *>(dz2 + row * ldz2 + c));
            g.x += g2.x; g.y += g2.y; g.z += g2.z; g.w += g2.w;
        }
        return g;
    };
    auto consume = [&](int rr, const float4 &g4) {
        const float2 *xs = reinterpret_cast<const float2 *>(s_x + rr * 10);
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

    for (unsigned r0 = blockIdx.x * (unsigned)TILE; r0 < rpc; r0 += gridDim.x * (unsigned)TILE) {
        stage_tile(s_x, xb, ib, r0, rpc, K, shiftK, xbar);
        const int nrows = (int)min((unsigned)TILE, rpc - r0);
        const size_t row0 = cloud_row0 + r0;
        // the gradient rows of this tile do not depend on the staging: request the first pair before the barrier
        int rr = rl;
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
        if (rr < nrows) g0 = load_g(row0 + rr);
        if (rr + rpb < nrows) g1 = load_g(row0 + rr + rpb);
        __syncthreads();
        for (; rr + rpb < nrows; rr += 2 * rpb) {
            float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f), n1 = n0;
            if (rr + 2 * rpb < nrows) n0 = load_g(row0 + rr + 2 * rpb);
            if (rr + 3 * rpb < nrows) n1 = load_g(row0 + rr + 3 * rpb);
            consume(rr, g0);
            consume(rr + rpb, g1);
            g0 = n0; g1 = n1;
        }
        if (rr < nrows) consume(rr, g0);
        __syncthreads();
    }

    // CTA reduction over the threads that share a column group, fixed order: lanes (xor butterfly) -> warps (ascending)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = cq; o < 32; o <<= 1) {
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[a][k] += __shfl_xor_sync(0xffffffffu, acc[a][k], o);
    }
    const int owners = cq < 32 ? cq : 32;          // lanes of a warp holding distinct column groups
    const int wpg = cq < 32 ? 1 : cq / 32;         // warps needed to cover all column groups once
    float *pb = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * NACC * h;
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

int pu_locse_moments(const float *xyz, const int32_t *idx, int B, int N, int K, float *mom, void *workspace,
                     size_t workspace_bytes, pu_stream_t stream) {
    if (!xyz || !idx || !mom) return PU_ERR_INVALID_ARG;
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
    locse_moment_kernel<0><<<grid, 256, 0, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, nullptr, 0.f, part);
    PU_LAUNCH_CHECK();
    launch_reduce_parts(part, chunks, 10, mom, 0, st);
    PU_LAUNCH_CHECK();
    locse_moment_kernel<1><<<grid, 256, 0, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, mom, inv_count, part);
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

int pu_locse_mlp_fwd(const float *xyz, const int32_t *idx, int B, int N, int K, const float *w, int h, const float *coef,
                     float slope, float *out, int ldo, float *out2, int ldo2, pu_stream_t stream) {
    if (!xyz || !idx || !w || !coef || !out || ldo < h || (out2 && ldo2 < h)) return PU_ERR_INVALID_ARG;
    if (!width_ok(h)) return PU_ERR_UNSUPPORTED;
    if ((((uintptr_t)out) & 15) || (ldo & 3) || (out2 && ((((uintptr_t)out2) & 15) || (ldo2 & 3))) || (((uintptr_t)w) & 15) ||
        (((uintptr_t)coef) & 15))
        return PU_ERR_INVALID_ARG;
    Geom g;
    const int rc = geom(B, N, K, &g);
    if (rc != PU_OK) return rc;
    if (B == 0 || g.rpc == 0) return PU_OK;
    const dim3 grid(grid_x(g.rpc, B, kNumSMs * 8), (unsigned)B);
    locse_mlp_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, idx, N, K, g.shiftK, g.rpc, w, h, coef, slope, out, ldo, out2,
                                                                 ldo2);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_locse_mlp_bwd(const float *xyz, const int32_t *idx, int B, int N, int K, const float *w, int h, const float *coef,
                     const float *gamma, const float *bias, int training, float slope, const float *dz, int ldz, const float *dz2, int ldz2,
                     float *dw, int accumulate_dw, float *dbias, float *dgamma, float *dbeta, void *workspace,
                     size_t workspace_bytes, pu_stream_t stream) {
    if (!xyz || !idx || !w || !coef || !gamma || !bias || !dz || !dw || !dgamma || !dbeta || ldz < h || (dz2 && ldz2 < h))
        return PU_ERR_INVALID_ARG;
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
    const dim3 grid(grid_x(g.rpc, B, BWD_CTAS), (unsigned)B);
    locse_mlp_bwd_kernel<<<grid, 256, 0, st>>>(xyz, idx, N, K, g.shiftK, g.rpc, w, h, coef, slope, dz, ldz, dz2, ldz2, part);
    PU_LAUNCH_CHECK();
    locse_mlp_bwd_finalize_kernel<<<h, 256, 0, st>>>(part, (int)(grid.x * grid.y), h, w, coef, gamma, bias, training ? 1 : 0, dw,
                                                     accumulate_dw, dbias, dgamma, dbeta);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

}  // extern "C"
