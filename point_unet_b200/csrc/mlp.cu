// mlp.cu -- the contraction kernels of the PointSegment hot path (fp32 CUDA-core tiles).
//
//   pu_linear_fwd          helper_tf_util.conv2d / conv2d_transpose 1x1 (PointSegment/helper_tf_util.py:115-170,
//                          173-250) and tf.layers.dense (RandLANet.py:114): y = x W + b, with the per-channel
//                          batch-norm statistics (sum, sum of squares over all rows) accumulated in the epilogue.
//   pu_att_pooling_fwd     Network.att_pooling (RandLANet.py:388-401) up to f_agg: FC (no bias), softmax over the
//                          K neighbours per channel, weighted sum -- ONE kernel, the [B,N,K,d] activations /
//                          scores / products of the reference never touch HBM.
//   pu_att_pooling_bwd     its gradient: recomputes the scores tile, emits d(act) and the direct term g*s.
//   pu_wgrad               dW = x^T dy (+ db), deterministic two-stage reduction over row chunks.
//   pu_bn_act_* / pu_act_* batch-norm apply + LeakyReLU(0.2) forward/backward (helper_tf_util.py:166-169),
//                          residual add of dilated_res_block (RandLANet.py:321).
//
// One tiled kernel template serves the three GEMM-shaped ops; they differ in the epilogue only.  Tile
// 128 rows x 64 channels, 256 threads, 8x4 register micro-tile; 16 consecutive rows = the K=16 neighbours of one
// point, held by two lanes (l, l^16) of one warp, so softmax over K is registers + one shuffle.
#include <float.h>

#include "common.cuh"

namespace pu {
namespace mlp {

enum { EPI_STORE = 0, EPI_ATT_FWD = 1, EPI_ATT_BWD = 2 };

struct GemmParams {
    const float *A; int lda;   // [M,K]
    const float *B; int ldb;   // [K,N]
    float *C; int ldc;         // [M,N]   (EPI_ATT_BWD: d(act))
    const float *bias;         // [N] or null
    long long M; int N, K;
    int accumulate;            // C += A B
    float *stat_sum, *stat_sq; // [gridDim.x, N] partial column sums of the stored value, or null
    const float *X; int ldx;   // att: feature_set rows [M,N]
    const float *G; int ldg;   // att bwd: upstream gradient [M/16, N]
    float *OUT; int ldo;       // att fwd: f_agg [M/16, N];  att bwd: g * s  [M,N]
    int c_bf16;                // narrow linear only: C is stored as bf16 (ldc in elements)
};

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN, int EPI>
__global__ void __launch_bounds__(256) gemm_kernel(const GemmParams p) {
    constexpr int TX = BN / TN, TY = BM / TM;
    static_assert(TX * TY == 256, "256 threads");
    static_assert(TN == 4 && (TM % 4) == 0, "micro-tile");
    constexpr int LDA_S = BM + 4;
    __shared__ __align__(16) float As[BK][LDA_S];
    __shared__ __align__(16) float Bs[BK][BN];

    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const bool vecA = ((p.lda & 3) == 0) && ((((uintptr_t)p.A) & 15) == 0);
    const bool vecB = ((p.ldb & 3) == 0) && ((((uintptr_t)p.B) & 15) == 0);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += BK) {
        // A tile: BM rows x BK, stored k-major
        for (int i = tid; i < BM * (BK / 4); i += 256) {
            const int row = i / (BK / 4), kq = (i % (BK / 4)) * 4;
            const long long gm = m0 + row;
            const int gk = k0 + kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gm < p.M) {
                const float *a = p.A + (size_t)gm * p.lda + gk;
                if (vecA && gk + 3 < p.K) {
                    v = *reinterpret_cast<const float4 *>(a);
                } else {
                    if (gk + 0 < p.K) v.x = a[0];
                    if (gk + 1 < p.K) v.y = a[1];
                    if (gk + 2 < p.K) v.z = a[2];
                    if (gk + 3 < p.K) v.w = a[3];
                }
            }
            As[kq + 0][row] = v.x; As[kq + 1][row] = v.y; As[kq + 2][row] = v.z; As[kq + 3][row] = v.w;
        }
        // B tile: BK x BN
        for (int i = tid; i < BK * (BN / 4); i += 256) {
            const int kr = i / (BN / 4), nq = (i % (BN / 4)) * 4;
            const int gk = k0 + kr, gn = n0 + nq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gk < p.K) {
                const float *b = p.B + (size_t)gk * p.ldb + gn;
                if (vecB && gn + 3 < p.N) {
                    v = *reinterpret_cast<const float4 *>(b);
                } else {
                    if (gn + 0 < p.N) v.x = b[0];
                    if (gn + 1 < p.N) v.y = b[1];
                    if (gn + 2 < p.N) v.z = b[2];
                    if (gn + 3 < p.N) v.w = b[3];
                }
            }
            *reinterpret_cast<float4 *>(&Bs[kr][nq]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(&As[k][ty * TM + i]);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
            {
                const float4 v = *reinterpret_cast<const float4 *>(&Bs[k][tx * TN]);
                b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    const int gn = n0 + tx * TN;
    if constexpr (EPI == EPI_STORE) {
        float cs[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) cs[j] = 0.f;
        float (&vals)[TM][TN] = acc;  // the stored values overwrite the accumulators in place
        float bj[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) bj[j] = (p.bias && gn + j < p.N) ? p.bias[gn + j] : 0.f;
        const bool vecC = ((p.ldc & 3) == 0) && ((((uintptr_t)p.C) & 15) == 0) && (gn + 3 < p.N);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const long long gm = m0 + ty * TM + i;
            if (gm >= p.M) continue;
            float *c = p.C + (size_t)gm * p.ldc + gn;
            float (&v)[TN] = vals[i];
#pragma unroll
            for (int j = 0; j < TN; ++j) v[j] = acc[i][j] + bj[j];
            if (vecC) {
                if (p.accumulate) {
                    const float4 o = *reinterpret_cast<const float4 *>(c);
                    v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
                }
                *reinterpret_cast<float4 *>(c) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (gn + j < p.N) {
                        if (p.accumulate) v[j] += c[j];
                        c[j] = v[j];
                    }
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) cs[j] += v[j];
        }
        if (p.stat_sum) {  // per-CTA column partials (fixed order => deterministic)
            // Robust batch statistics: each tile emits (sum, M2) with M2 = sum (v - tile_mean)^2 computed from the
            // values still in registers; pu_stats_finalize merges tiles with Chan's parallel-variance formula in
            // double.  (E[y^2] - E[y]^2 loses ~|mean|/sigma digits; this form does not.)
            float *red = &As[0][0];  // TY x BN floats <= BK*LDA_S
            static_assert(TY * BN <= BK * LDA_S, "reduction scratch");
            __shared__ float s_mean[BN];
            const long long rows_here = min((long long)BM, p.M - m0);
#pragma unroll
            for (int j = 0; j < TN; ++j) red[ty * BN + tx * TN + j] = cs[j];
            __syncthreads();
            if (tid < BN) {
                float s = 0.f;
                for (int r = 0; r < TY; ++r) s += red[r * BN + tid];
                s_mean[tid] = s / (float)rows_here;
                if (n0 + tid < p.N) p.stat_sum[(size_t)blockIdx.x * p.N + n0 + tid] = s;
            }
            __syncthreads();
            float m2[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) m2[j] = 0.f;
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                if (m0 + ty * TM + i >= p.M) continue;
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const float dlt = vals[i][j] - s_mean[tx * TN + j];
                    m2[j] = fmaf(dlt, dlt, m2[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) red[ty * BN + tx * TN + j] = m2[j];
            __syncthreads();
            if (tid < BN && n0 + tid < p.N) {
                float q = 0.f;
                for (int r = 0; r < TY; ++r) q += red[r * BN + tid];
                p.stat_sq[(size_t)blockIdx.x * p.N + n0 + tid] = q;
            }
        }
    } else {
        // A point = 16 consecutive rows.  <128,64,8,4>: a thread holds 8 of them, the other half lives in lane ^ 16.
        // <256,16,4,4>: a thread holds 4, the other three quarters live in lanes ^4, ^8, ^12 (ty = lane >> 2).
        static_assert((BM == 128 && TM == 8 && TX == 16) || (BM == 256 && TM == 4 && TX == 4), "att epilogue layout");
        auto point_reduce_max = [](float v) {
            if constexpr (TM == 8) return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
            else { v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4)); return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8)); }
        };
        auto point_reduce_sum = [](float v) {
            if constexpr (TM == 8) return v + __shfl_xor_sync(0xffffffffu, v, 16);
            else { v += __shfl_xor_sync(0xffffffffu, v, 4); return v + __shfl_xor_sync(0xffffffffu, v, 8); }
        };
        const long long pt = m0 / 16 + (ty * TM) / 16;
        const bool writer = ((ty * TM) & 15) == 0;
        const bool pt_ok = (m0 + ty * TM) < p.M && gn < p.N;
        float cmax[TN], csum[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float m = acc[0][j];
#pragma unroll
            for (int i = 1; i < TM; ++i) m = fmaxf(m, acc[i][j]);
            cmax[j] = point_reduce_max(m);
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < TM; ++i) { acc[i][j] = ex2_approx(fmaf(acc[i][j], 1.4426950408889634f, -cmax[j] * 1.4426950408889634f)); s += acc[i][j]; }
            csum[j] = point_reduce_sum(s);
        }
        float x[TM][TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pt_ok) v = *reinterpret_cast<const float4 *>(p.X + (size_t)(m0 + ty * TM + i) * p.ldx + gn);
            x[i][0] = v.x; x[i][1] = v.y; x[i][2] = v.z; x[i][3] = v.w;
        }
        if constexpr (EPI == EPI_ATT_FWD) {
            float num[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < TM; ++i) s = fmaf(x[i][j], acc[i][j], s);
                num[j] = point_reduce_sum(s);
            }
            if (pt_ok && writer)
                *reinterpret_cast<float4 *>(p.OUT + (size_t)pt * p.ldo + gn) =
                    make_float4(num[0] * rcp_approx(csum[0]), num[1] * rcp_approx(csum[1]), num[2] * rcp_approx(csum[2]),
                                num[3] * rcp_approx(csum[3]));
        } else {  // EPI_ATT_BWD
            float g[TN] = {0.f, 0.f, 0.f, 0.f};
            if (pt_ok) {
                const float4 v = *reinterpret_cast<const float4 *>(p.G + (size_t)pt * p.ldg + gn);
                g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
            }
            float dot[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const float inv = rcp_approx(csum[j]);
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    acc[i][j] *= inv;                      // s_k  (softmax score)
                    s = fmaf(g[j] * x[i][j], acc[i][j], s);  // sum_k ds_k s_k,  ds_k = g x_k
                }
                dot[j] = point_reduce_sum(s);
            }
            if (pt_ok) {
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const size_t row = (size_t)(m0 + ty * TM + i);
                    float da[TN], dx[TN];
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        da[j] = acc[i][j] * (g[j] * x[i][j] - dot[j]);
                        dx[j] = g[j] * acc[i][j];
                    }
                    *reinterpret_cast<float4 *>(p.C + row * p.ldc + gn) = make_float4(da[0], da[1], da[2], da[3]);
                    *reinterpret_cast<float4 *>(p.OUT + row * p.ldo + gn) = make_float4(dx[0], dx[1], dx[2], dx[3]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// narrow linear (K <= 16, N <= 32): HBM-bound.  Thread (row lane, y) owns the 4 output columns 4y..4y+3 with the K x 4
// weight slice in REGISTERS and walks the rows of its CTA's chunk: one vector load of the x row, 4K FMAs, one 128-bit
// store.  Each CTA covers exactly NARROW_ROWS consecutive rows (= one statistics tile); batch-norm partials (sum, centred
// M2, shifted by the tile's first output row) are reduced per CTA in fixed order.
constexpr int NARROW_ROWS = 2048;
template <int KP>
__global__ void __launch_bounds__(256, 3) linear_narrow_kernel(const GemmParams p) {
    __shared__ float s_red[8 * 32 * 2];            // [warp][NY*4 columns][2]
    __shared__ __align__(16) float s_w[KP * 32];   // weight [k][32 columns], zero padded (broadcast reads)
    int NY = 1;
    while (NY * 4 < p.N) NY <<= 1;           // column groups of 4 (power of two <= 8)
    const int RL = 256 / NY;
    const int y = threadIdx.x % NY, rl = threadIdx.x / NY;
    const int c0 = y * 4;
    const long long r_begin = (long long)blockIdx.x * NARROW_ROWS;
    const long long r_end = min(p.M, r_begin + NARROW_ROWS);
    for (int i = threadIdx.x; i < KP * 32; i += 256) {
        const int k = i >> 5, c = i & 31;
        s_w[i] = (k < p.K && c < p.N) ? p.B[(size_t)k * p.ldb + c] : 0.f;
    }
    __syncthreads();
    float bj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bj[j] = (p.bias && c0 + j < p.N) ? p.bias[c0 + j] : 0.f;
    const bool va4 = ((p.lda & 3) == 0) && ((((uintptr_t)p.A) & 15) == 0) && ((p.K & 3) == 0);
    const bool va2 = ((p.lda & 1) == 0) && ((((uintptr_t)p.A) & 7) == 0) && ((p.K & 1) == 0);
    const bool vc4 = ((p.ldc & 3) == 0) && ((((uintptr_t)p.C) & (p.c_bf16 ? 7 : 15)) == 0) && (c0 + 3 < p.N);
    auto load_row = [&](long long r, float (&xv)[KP]) {
        const float *a = p.A + (size_t)r * p.lda;
        if (va4) {
#pragma unroll
            for (int i = 0; i < KP; i += 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < p.K) v = *reinterpret_cast<const float4 *>(a + i);
                xv[i] = v.x; xv[i + 1] = v.y; xv[i + 2] = v.z; xv[i + 3] = v.w;
            }
        } else if (va2) {
#pragma unroll
            for (int i = 0; i < KP; i += 2) {
                float2 v = make_float2(0.f, 0.f);
                if (i < p.K) v = *reinterpret_cast<const float2 *>(a + i);
                xv[i] = v.x; xv[i + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < KP; ++i) xv[i] = i < p.K ? a[i] : 0.f;
        }
    };
    auto dot_row = [&](const float (&xv)[KP], float (&o)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = bj[j];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const float4 wk = *reinterpret_cast<const float4 *>(&s_w[k * 32 + c0]);
            o[0] = fmaf(xv[k], wk.x, o[0]); o[1] = fmaf(xv[k], wk.y, o[1]);
            o[2] = fmaf(xv[k], wk.z, o[2]); o[3] = fmaf(xv[k], wk.w, o[3]);
        }
    };
    float sh[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.stat_sum && r_begin < r_end) {  // shift = output of the tile's first row (every thread of a column group agrees)
        float xv[KP];
        load_row(r_begin, xv);
        dot_row(xv, sh);
    }
    auto finish_row = [&](long long r, float (&o)[4], const float4 &old) {
        float *c = p.C + (size_t)r * p.ldc + c0;
        if (p.c_bf16) {   // bf16 storage of the pre-normalisation activations; the statistics below stay on the fp32 values
            unsigned short *cb = reinterpret_cast<unsigned short *>(p.C) + (size_t)r * p.ldc + c0;
            if (vc4) {
                *reinterpret_cast<uint2 *>(cb) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c0 + j < p.N) cb[j] = (unsigned short)(pack_bf16x2(o[j], 0.f) & 0xffffu);
            }
        } else if (vc4) {
            if (p.accumulate) { o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w; }
            *reinterpret_cast<float4 *>(c) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < p.N) {
                    if (p.accumulate) o[j] += c[j];
                    c[j] = o[j];
                }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float d = o[j] - sh[j]; s1[j] += d; s2[j] = fmaf(d, d, s2[j]); }
    };
    long long r = r_begin + rl;
    for (; r + RL < r_end; r += 2 * RL) {  // two rows per trip: both x rows (and old C values) are in flight together
        float xa[KP], xb[KP], oa[4], ob[4];
        load_row(r, xa);
        load_row(r + RL, xb);
        float4 olda = make_float4(0.f, 0.f, 0.f, 0.f), oldb = olda;
        if (p.accumulate && vc4) {
            olda = *reinterpret_cast<const float4 *>(p.C + (size_t)r * p.ldc + c0);
            oldb = *reinterpret_cast<const float4 *>(p.C + (size_t)(r + RL) * p.ldc + c0);
        }
        dot_row(xa, oa);
        dot_row(xb, ob);
        finish_row(r, oa, olda);
        finish_row(r + RL, ob, oldb);
    }
    for (; r < r_end; r += RL) {
        float xa[KP], oa[4];
        load_row(r, xa);
        float4 olda = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.accumulate && vc4) olda = *reinterpret_cast<const float4 *>(p.C + (size_t)r * p.ldc + c0);
        dot_row(xa, oa);
        finish_row(r, oa, olda);
    }
    if (p.stat_sum) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int o = NY; o < 32; o <<= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
            }
        }
        if (lane < NY) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s_red[((warp * 32) + y * 4 + j) * 2 + 0] = s1[j];
                s_red[((warp * 32) + y * 4 + j) * 2 + 1] = s2[j];
            }
        }
        __syncthreads();
        if (threadIdx.x < NY * 4 && threadIdx.x < p.N) {
            float a = 0.f, b = 0.f;
            for (int wv = 0; wv < 8; ++wv) { a += s_red[(wv * 32 + threadIdx.x) * 2]; b += s_red[(wv * 32 + threadIdx.x) * 2 + 1]; }
            // the shift of column threadIdx.x lives in the thread with y = threadIdx.x / 4: recompute it here
            float shc = p.bias ? p.bias[threadIdx.x] : 0.f;
            const float *a0 = p.A + (size_t)r_begin * p.lda;
            for (int k = 0; k < p.K; ++k) shc = fmaf(a0[k], p.B[(size_t)k * p.ldb + threadIdx.x], shc);
            const float n = (float)(r_end - r_begin);
            p.stat_sum[(size_t)blockIdx.x * p.N + threadIdx.x] = fmaf(n, shc, a);
            p.stat_sq[(size_t)blockIdx.x * p.N + threadIdx.x] = fmaxf(b - a * a / n, 0.f);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// wgrad: dW[K,N] = sum_r A[r,K]^T G[r,N]  (+ db[N] = sum_r G[r,N]); grid (k tiles, n tiles, row chunks);
// each CTA writes its partial tile to part[chunk][K][N]; pu_wgrad then reduces over chunks in a fixed order.
constexpr int WT = 64, WR = 16;
__global__ void __launch_bounds__(256) wgrad_kernel(const float *__restrict__ A, int lda, const float *__restrict__ G,
                                                    int ldg, long long M, int K, int N, long long rows_per_chunk,
                                                    float *__restrict__ part, float *__restrict__ db_part) {
    __shared__ __align__(16) float As[WR][WT];
    __shared__ __align__(16) float Gs[WR][WT];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int k0 = blockIdx.x * WT, n0 = blockIdx.y * WT;
    const long long r_begin = (long long)blockIdx.z * rows_per_chunk;
    const long long r_end = min(M, r_begin + rows_per_chunk);
    const bool vecA = ((lda & 3) == 0) && ((((uintptr_t)A) & 15) == 0);
    const bool vecG = ((ldg & 3) == 0) && ((((uintptr_t)G) & 15) == 0);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float dbacc[4] = {0.f, 0.f, 0.f, 0.f};

    for (long long r0 = r_begin; r0 < r_end; r0 += WR) {
        {   // 16 rows x 64 cols of each operand = 256 float4, one per thread
            const int rr = tid / 16, cq = (tid % 16) * 4;
            const long long gr = r0 + rr;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vg = va;
            if (gr < r_end) {
                const float *a = A + (size_t)gr * lda + k0 + cq;
                if (vecA && k0 + cq + 3 < K) va = *reinterpret_cast<const float4 *>(a);
                else {
                    if (k0 + cq + 0 < K) va.x = a[0];
                    if (k0 + cq + 1 < K) va.y = a[1];
                    if (k0 + cq + 2 < K) va.z = a[2];
                    if (k0 + cq + 3 < K) va.w = a[3];
                }
                const float *g = G + (size_t)gr * ldg + n0 + cq;
                if (vecG && n0 + cq + 3 < N) vg = *reinterpret_cast<const float4 *>(g);
                else {
                    if (n0 + cq + 0 < N) vg.x = g[0];
                    if (n0 + cq + 1 < N) vg.y = g[1];
                    if (n0 + cq + 2 < N) vg.z = g[2];
                    if (n0 + cq + 3 < N) vg.w = g[3];
                }
            }
            *reinterpret_cast<float4 *>(&As[rr][cq]) = va;
            *reinterpret_cast<float4 *>(&Gs[rr][cq]) = vg;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < WR; ++r) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[r][ty * 4]);
            const float4 g = *reinterpret_cast<const float4 *>(&Gs[r][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
            if (ty == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) dbacc[j] += gv[j];
            }
        }
        __syncthreads();
    }
    float *o = part + (size_t)blockIdx.z * K * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gk = k0 + ty * 4 + i;
        if (gk >= K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn < N) o[(size_t)gk * N + gn] = acc[i][j];
        }
    }
    if (db_part && ty == 0 && blockIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn < N) db_part[(size_t)blockIdx.z * N + gn] = dbacc[j];
        }
    }
}

// mean/var (biased) per channel from per-tile (sum, M2) partials: Chan et al. parallel merge, in double.
//   var = [ sum_t M2_t + sum_t n_t mean_t^2 - n mean^2 ] / n      (the subtraction is done in double)
// One CTA per channel, 1024 threads stride over the tiles (the tensor-core kernels emit one partial per 128 rows: 22 500 of
// them for the 2.9 M-row tensors, so the per-thread chains must be short), fixed-order block reduction.
constexpr int SF_THREADS = 1024;
__global__ void __launch_bounds__(SF_THREADS) stats_finalize_kernel(const float *__restrict__ psum, const float *__restrict__ pm2,
                                                                    int tiles, int C, long long count, int rows_per_tile,
                                                                    float *__restrict__ mean, float *__restrict__ var,
                                                                    const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                    float eps, float *__restrict__ invstd, float *__restrict__ scale,
                                                                    float *__restrict__ shift, float *__restrict__ moving_mean,
                                                                    float *__restrict__ moving_var, float momentum, float unbias) {
    constexpr int NT = SF_THREADS;
    __shared__ double red[3][NT];
    const int c = blockIdx.x;
    double s = 0.0, q = 0.0, r = 0.0;
    int t = threadIdx.x;
    for (; t + 3 * NT < tiles; t += 4 * NT) {  // four tiles per trip: eight independent loads in flight per thread
        float st[4], mt[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            st[j] = psum[(size_t)(t + NT * j) * C + c];
            mt[j] = pm2[(size_t)(t + NT * j) * C + c];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long r0 = (long long)(t + NT * j) * rows_per_tile;
            const double nt = (double)min((long long)rows_per_tile, count - r0);
            s += (double)st[j];
            q += (double)mt[j];
            r += (double)st[j] * (double)st[j] / nt;
        }
    }
    for (; t < tiles; t += NT) {
        const long long r0 = (long long)t * rows_per_tile;
        const double nt = (double)min((long long)rows_per_tile, count - r0);
        const double st = (double)psum[(size_t)t * C + c];
        s += st;
        q += (double)pm2[(size_t)t * C + c];
        r += st * st / nt;
    }
    red[0][threadIdx.x] = s; red[1][threadIdx.x] = q; red[2][threadIdx.x] = r;
    __syncthreads();
    for (int o = NT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            red[0][threadIdx.x] += red[0][threadIdx.x + o];
            red[1][threadIdx.x] += red[1][threadIdx.x + o];
            red[2][threadIdx.x] += red[2][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double n = (double)count, m = red[0][0] / n;
        const double v = (red[1][0] + red[2][0] - n * m * m) / n;
        const float mf = (float)m, vf = (float)(v > 0.0 ? v : 0.0);
        mean[c] = mf;
        var[c] = vf;
        if (gamma) {   // fused pu_bn_prepare (same arithmetic as bn_prepare_kernel): one launch per batch norm instead of two
            const float is = rsqrtf(vf + eps);
            invstd[c] = is;
            scale[c] = gamma[c] * is;
            shift[c] = mf;
            shift[C + c] = beta[c];
            if (moving_mean) {
                moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * mf;
                moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * vf * unbias;
            }
        }
    }
}

// The same for SHORT partial lists (the deep levels: 3-350 tiles, but 256-1024 channels): one WARP per channel, eight channels
// per CTA, lanes stride over the tiles, xor-shuffle reduction in double.  One 1024-thread CTA per channel spent its time in
// the ten barrier rounds of the block reduction with most threads idle: 28 us per launch at 1024 channels (ncu launch list),
// between a conv and its batch norm on the critical path.  Same arithmetic; the summation order depends only on `tiles`.
__global__ void __launch_bounds__(256) stats_finalize_warp_kernel(const float *__restrict__ psum, const float *__restrict__ pm2,
                                                                  int tiles, int C, long long count, int rows_per_tile,
                                                                  float *__restrict__ mean, float *__restrict__ var,
                                                                  const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                  float eps, float *__restrict__ invstd, float *__restrict__ scale,
                                                                  float *__restrict__ shift, float *__restrict__ moving_mean,
                                                                  float *__restrict__ moving_var, float momentum, float unbias) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    double s = 0.0, q = 0.0, r = 0.0;
    for (int t = lane; t < tiles; t += 32) {
        const long long r0 = (long long)t * rows_per_tile;
        const double nt = (double)min((long long)rows_per_tile, count - r0);
        const double st = (double)psum[(size_t)t * C + c];
        s += st;
        q += (double)pm2[(size_t)t * C + c];
        r += st * st / nt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
        r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    if (lane == 0) {
        const double n = (double)count, m = s / n;
        const double v = (q + r - n * m * m) / n;
        const float mf = (float)m, vf = (float)(v > 0.0 ? v : 0.0);
        mean[c] = mf;
        var[c] = vf;
        if (gamma) {
            const float is = rsqrtf(vf + eps);
            invstd[c] = is;
            scale[c] = gamma[c] * is;
            shift[c] = mf;
            shift[C + c] = beta[c];
            if (moving_mean) {
                moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * mf;
                moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * vf * unbias;
            }
        }
    }
}
constexpr int SF_WARP_MAX_TILES = 64;   // two rounds per lane; longer lists: the 1024-thread kernel (measured: 15 us here vs 7 us there at 352 tiles)

// NOTE on "shift": all batch-norm kernels below evaluate z = (y - mean[c]) * scale[c] + beta[c] (centered form, no
// cancellation between y*scale and mean*scale); `shift` is a [2,C] array: row 0 = mean, row 1 = beta.
__device__ __forceinline__ float4 bn_z(const float4 v, const float4 mu, const float4 sc, const float4 be) {
    return make_float4(fmaf(v.x - mu.x, sc.x, be.x), fmaf(v.y - mu.y, sc.y, be.y), fmaf(v.z - mu.z, sc.z, be.z),
                       fmaf(v.w - mu.w, sc.w, be.w));
}

__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const void *__restrict__ y, int ld_y, const float *__restrict__ scale,
                                                         const float *__restrict__ shift, const void *__restrict__ y2,
                                                         int ld_y2, const float *__restrict__ scale2,
                                                         const float *__restrict__ shift2, float slope, long long R, int C,
                                                         float *__restrict__ out, int ld_o, float *__restrict__ out2, int ld_o2,
                                                         int y_bf16, int y2_bf16) {
    // a thread owns ONE float4 column group for its lifetime (coefficients in registers, no index division in the loop)
    // and walks the rows; consecutive threads still touch consecutive 16-byte chunks
    const int cq = C >> 2, rpb = 256 / cq;
    const int c = (threadIdx.x % cq) * 4, rl = threadIdx.x / cq;
    if (rl >= rpb) return;
    const float4 sc = *reinterpret_cast<const float4 *>(scale + c), mu = *reinterpret_cast<const float4 *>(shift + c),
                 be = *reinterpret_cast<const float4 *>(shift + C + c);
    float4 s2 = sc, m2 = mu, b2 = be;
    if (y2) {
        s2 = *reinterpret_cast<const float4 *>(scale2 + c);
        m2 = *reinterpret_cast<const float4 *>(shift2 + c);
        b2 = *reinterpret_cast<const float4 *>(shift2 + C + c);
    }
    auto emit = [&](long long r, const float4 &v, const float4 &w) {
        const float4 z1 = bn_z(v, mu, sc, be);
        float z[4] = {z1.x, z1.y, z1.z, z1.w};
        if (y2) {
            const float4 z2 = bn_z(w, m2, s2, b2);
            z[0] += z2.x; z[1] += z2.y; z[2] += z2.z; z[3] += z2.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = z[j] > 0.f ? z[j] : z[j] * slope;
        *reinterpret_cast<float4 *>(out + (size_t)r * ld_o + c) = make_float4(z[0], z[1], z[2], z[3]);
        if (out2) *reinterpret_cast<float4 *>(out2 + (size_t)r * ld_o2 + c) = make_float4(z[0], z[1], z[2], z[3]);
    };
    const long long step = (long long)gridDim.x * rpb;
    long long r = (long long)blockIdx.x * rpb + rl;
    for (; r + step < R; r += 2 * step) {  // two rows in flight
        const float4 v0 = load4_any(y, y_bf16, (size_t)r, ld_y, c);
        const float4 v1 = load4_any(y, y_bf16, (size_t)(r + step), ld_y, c);
        float4 w0 = v0, w1 = v1;
        if (y2) {
            w0 = load4_any(y2, y2_bf16, (size_t)r, ld_y2, c);
            w1 = load4_any(y2, y2_bf16, (size_t)(r + step), ld_y2, c);
        }
        emit(r, v0, w0);
        emit(r + step, v1, w1);
    }
    if (r < R) {
        const float4 v0 = load4_any(y, y_bf16, (size_t)r, ld_y, c);
        float4 w0 = v0;
        if (y2) w0 = load4_any(y2, y2_bf16, (size_t)r, ld_y2, c);
        emit(r, v0, w0);
    }
}

// dz = dout * (out > 0 ? 1 : slope)   -- gradient through the LeakyReLU given its OUTPUT (sign-preserving)
__global__ void __launch_bounds__(256) act_bwd_kernel(const float *__restrict__ dout, int ld_d, const float *__restrict__ out,
                                                      int ld_o, float slope, long long R, int C, float *__restrict__ dz,
                                                      int ld_z) {
    const int cq = C >> 2;
    const long long total = R * cq;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / cq;
        const int c = (int)(t - r * cq) * 4;
        const float4 g = *reinterpret_cast<const float4 *>(dout + (size_t)r * ld_d + c);
        const float4 o = *reinterpret_cast<const float4 *>(out + (size_t)r * ld_o + c);
        *reinterpret_cast<float4 *>(dz + (size_t)r * ld_z + c) =
            make_float4(o.x > 0.f ? g.x : g.x * slope, o.y > 0.f ? g.y : g.y * slope, o.z > 0.f ? g.z : g.z * slope,
                        o.w > 0.f ? g.w : g.w * slope);
    }
}

// BN backward, pass 1: per-channel partials of  sum(dz)  and  sum(dz * y)  with dz = dout * lrelu'(y*scale+shift).
// grid = fixed number of CTAs; each CTA strides over rows; thread owns one float4 column group.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float *__restrict__ dout, int ld_d,
                                                            const float *__restrict__ dout2, int ld_d2,
                                                            const void *__restrict__ y, int ld_y,
                                                            const float *__restrict__ scale, const float *__restrict__ shift,
                                                            float slope, long long R, int C, float *__restrict__ part_dz,
                                                            float *__restrict__ part_dzy, int y_bf16) {
    extern __shared__ float sred[];  // [rows_in_flight][C][2]
    const int cq = C >> 2;                 // column groups
    const int rpb = 256 / cq > 0 ? 256 / cq : 1;  // rows processed concurrently by the CTA
    const int my_c = (threadIdx.x % cq) * 4, my_r = threadIdx.x / cq;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (my_r < rpb) {
        const float4 sc = *reinterpret_cast<const float4 *>(scale + my_c), mu = *reinterpret_cast<const float4 *>(shift + my_c),
                     be = *reinterpret_cast<const float4 *>(shift + C + my_c);
        for (long long r = (long long)blockIdx.x * rpb + my_r; r < R; r += (long long)gridDim.x * rpb) {
            float4 g = *reinterpret_cast<const float4 *>(dout + (size_t)r * ld_d + my_c);
            if (dout2) {
                const float4 g2 = *reinterpret_cast<const float4 *>(dout2 + (size_t)r * ld_d2 + my_c);
                g.x += g2.x; g.y += g2.y; g.z += g2.z; g.w += g2.w;
            }
            const float4 v = load4_any(y, y_bf16, (size_t)r, ld_y, my_c);
            const float4 z = bn_z(v, mu, sc, be);
            const float gz[4] = {z.x > 0.f ? g.x : g.x * slope, z.y > 0.f ? g.y : g.y * slope,
                                 z.z > 0.f ? g.z : g.z * slope, z.w > 0.f ? g.w : g.w * slope};
            const float yv[4] = {v.x - mu.x, v.y - mu.y, v.z - mu.z, v.w - mu.w};  // centered: sum dz*(y-mean)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[j] += gz[j]; q[j] = fmaf(gz[j], yv[j], q[j]); }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sred[((size_t)my_r * C + my_c + j) * 2 + 0] = s[j];
            sred[((size_t)my_r * C + my_c + j) * 2 + 1] = q[j];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = 0.f, b = 0.f;
        for (int r = 0; r < rpb; ++r) { a += sred[((size_t)r * C + c) * 2]; b += sred[((size_t)r * C + c) * 2 + 1]; }
        part_dz[(size_t)blockIdx.x * C + c] = a;
        part_dzy[(size_t)blockIdx.x * C + c] = b;
    }
}

// BN backward, pass 2: dy = ka[c]*dz + kb[c] + kc[c]*y
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float *__restrict__ dout, int ld_d,
                                                           const float *__restrict__ dout2, int ld_d2,
                                                           const void *__restrict__ y,
                                                           int ld_y, const float *__restrict__ scale,
                                                           const float *__restrict__ shift, float slope,
                                                           const float *__restrict__ ka, const float *__restrict__ kb,
                                                           const float *__restrict__ kc, long long R, int C,
                                                           float *__restrict__ dy, int ld_dy, int y_bf16) {
    // fixed column group per thread (coefficients in registers, no index division), two rows in flight
    const int cq = C >> 2, rpb = 256 / cq;
    const int c = (threadIdx.x % cq) * 4, rl = threadIdx.x / cq;
    if (rl >= rpb) return;
    const float4 sc = *reinterpret_cast<const float4 *>(scale + c), mu = *reinterpret_cast<const float4 *>(shift + c),
                 be = *reinterpret_cast<const float4 *>(shift + C + c);
    const float4 a = *reinterpret_cast<const float4 *>(ka + c), b = *reinterpret_cast<const float4 *>(kb + c),
                 cc = *reinterpret_cast<const float4 *>(kc + c);
    auto load_g = [&](long long r) {
        float4 g = *reinterpret_cast<const float4 *>(dout + (size_t)r * ld_d + c);
        if (dout2) {
            const float4 g2 = *reinterpret_cast<const float4 *>(dout2 + (size_t)r * ld_d2 + c);
            g.x += g2.x; g.y += g2.y; g.z += g2.z; g.w += g2.w;
        }
        return g;
    };
    auto emit = [&](long long r, const float4 &g, const float4 &v) {
        const float4 z = bn_z(v, mu, sc, be);
        const float gz[4] = {z.x > 0.f ? g.x : g.x * slope, z.y > 0.f ? g.y : g.y * slope,
                             z.z > 0.f ? g.z : g.z * slope, z.w > 0.f ? g.w : g.w * slope};
        // dy = ka*dz + kb + kc*(y - mean)
        *reinterpret_cast<float4 *>(dy + (size_t)r * ld_dy + c) =
            make_float4(fmaf(a.x, gz[0], fmaf(cc.x, v.x - mu.x, b.x)), fmaf(a.y, gz[1], fmaf(cc.y, v.y - mu.y, b.y)),
                        fmaf(a.z, gz[2], fmaf(cc.z, v.z - mu.z, b.z)), fmaf(a.w, gz[3], fmaf(cc.w, v.w - mu.w, b.w)));
    };
    const long long step = (long long)gridDim.x * rpb;
    long long r = (long long)blockIdx.x * rpb + rl;
    for (; r + step < R; r += 2 * step) {
        const float4 g0 = load_g(r), g1 = load_g(r + step);
        const float4 v0 = load4_any(y, y_bf16, (size_t)r, ld_y, c);
        const float4 v1 = load4_any(y, y_bf16, (size_t)(r + step), ld_y, c);
        emit(r, g0, v0);
        emit(r + step, g1, v1);
    }
    if (r < R) {
        const float4 g0 = load_g(r);
        const float4 v0 = load4_any(y, y_bf16, (size_t)r, ld_y, c);
        emit(r, g0, v0);
    }
}


// ---------------------------------------------------------------------------------------------
// narrow wgrad (K <= 16, N <= 32): HBM-bound.  Thread (row lane, y) owns rows r = lane, lane+RL, ... of the CTA's
// chunk and the 4 output columns 4y..4y+3; it keeps a KP x 4 accumulator block in REGISTERS, so every x / dy
// element is read once and costs one FMA -- no shared-memory traffic in the loop.  One block-level reduction per
// CTA at the end, then the usual fixed-order reduction over CTAs.
template <int KP>
__global__ void __launch_bounds__(256, KP > 8 ? 2 : 4) wgrad_narrow_kernel(const float *__restrict__ A, int lda,
                                                           const float *__restrict__ G, int ldg, long long M, int K,
                                                           int N, long long rows_per_chunk, float *__restrict__ part,
                                                           float *__restrict__ db_part) {
    __shared__ float s_red[8 * (KP + 1) * 8 * 4];  // [warp][(KP+1) rows of NY*4 columns]
    int NY = 1;                               // column groups of 4, rounded up to a power of two (<= 8)
    while (NY * 4 < min(N - (int)blockIdx.y * 32, 32)) NY <<= 1;
    const int RL = 256 / NY;                  // row lanes per CTA
    const int y = threadIdx.x % NY, rl = threadIdx.x / NY;
    const long long r_begin = (long long)blockIdx.x * rows_per_chunk;
    const long long r_end = min(M, r_begin + rows_per_chunk);
    // slab of the output this CTA owns: 32 columns (blockIdx.y) x 16 weight rows (blockIdx.z).  Wider layers are cut into
    // slabs along ONE axis (the launcher takes this path only when K <= 16 or N <= 32), so the wide operand is still read
    // once and only the thin one is re-read per slab.
    const int n0 = blockIdx.y * 32, k0 = blockIdx.z * 16;
    const int Kfull = K, Nfull = N;
    A += k0; G += n0;
    K = min(K - k0, 16); N = min(N - n0, 32);
    float acc[KP][4];
#pragma unroll
    for (int i = 0; i < KP; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool va4 = ((lda & 3) == 0) && ((((uintptr_t)A) & 15) == 0) && ((K & 3) == 0);
    const bool va2 = ((lda & 1) == 0) && ((((uintptr_t)A) & 7) == 0) && ((K & 1) == 0);
    const bool vg4 = ((ldg & 3) == 0) && ((((uintptr_t)G) & 15) == 0) && (y * 4 + 3 < N);
    auto load_row = [&](long long r, float (&av)[KP], float (&gv)[4]) {
        const float *a = A + (size_t)r * lda;
        const float *g = G + (size_t)r * ldg + y * 4;
        if (va4) {
#pragma unroll
            for (int i = 0; i < KP; i += 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < K) v = *reinterpret_cast<const float4 *>(a + i);
                av[i] = v.x; av[i + 1] = v.y; av[i + 2] = v.z; av[i + 3] = v.w;
            }
        } else if (va2) {
#pragma unroll
            for (int i = 0; i < KP; i += 2) {
                float2 v = make_float2(0.f, 0.f);
                if (i < K) v = *reinterpret_cast<const float2 *>(a + i);
                av[i] = v.x; av[i + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < KP; ++i) av[i] = i < K ? a[i] : 0.f;
        }
        if (vg4) {
            const float4 v = *reinterpret_cast<const float4 *>(g);
            gv[0] = v.x; gv[1] = v.y; gv[2] = v.z; gv[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) gv[j] = (y * 4 + j < N) ? g[j] : 0.f;
        }
    };
    auto accumulate = [&](const float (&av)[KP], const float (&gv)[4]) {
#pragma unroll
        for (int i = 0; i < KP; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) dbacc[j] += gv[j];
    };
    {
        // two rows per trip: all loads of both rows are issued before the first FMA (the row order of the sums is unchanged)
        long long r = r_begin + rl;
        if constexpr (KP > 8) {  // (the 8-wide variant already runs at 84 % of the copy bandwidth with one row per trip)
            for (; r + RL < r_end; r += 2 * RL) {
                float av0[KP], gv0[4], av1[KP], gv1[4];
                load_row(r, av0, gv0);
                load_row(r + RL, av1, gv1);
                accumulate(av0, gv0);
                accumulate(av1, gv1);
            }
        } else {
            for (; r + RL < r_end; r += RL) {
                float av0[KP], gv0[4];
                load_row(r, av0, gv0);
                accumulate(av0, gv0);
            }
        }
        if (r < r_end) {
            float av0[KP], gv0[4];
            load_row(r, av0, gv0);
            accumulate(av0, gv0);
        }
    }
    // reduction over row lanes: butterfly over the lanes of a warp that share y, then 8 warps through shared memory
    // (fixed order => deterministic)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (KP + 1) * NY * 4;  // floats per warp (acc rows + db row)
    for (int o = NY; o < 32; o <<= 1) {
#pragma unroll
        for (int i = 0; i < KP; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], o);
#pragma unroll
        for (int j = 0; j < 4; ++j) dbacc[j] += __shfl_xor_sync(0xffffffffu, dbacc[j], o);
    }
    if (lane < NY) {
        float *o = s_red + (size_t)warp * per;
#pragma unroll
        for (int i = 0; i < KP; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[(i * NY + y) * 4 + j] = acc[i][j];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[(KP * NY + y) * 4 + j] = dbacc[j];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < per; e += 256) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += s_red[(size_t)w * per + e];
        const int i = e / (NY * 4), col = e % (NY * 4);
        if (col < N) {
            if (i < K) part[(size_t)blockIdx.x * Kfull * Nfull + (size_t)(k0 + i) * Nfull + n0 + col] = s;
            else if (i == KP && db_part && k0 == 0) db_part[(size_t)blockIdx.x * Nfull + n0 + col] = s;
        }
    }
}

// per-channel batch-norm coefficients in one launch (replaces ~20 tiny elementwise launches per layer):
//   forward : invstd, scale = gamma*invstd, shift = beta - mean*scale; optional moving-average update
__global__ void bn_prepare_kernel(const float *__restrict__ mean, const float *__restrict__ var,
                                  const float *__restrict__ gamma, const float *__restrict__ beta, float eps, int C,
                                  float *__restrict__ invstd, float *__restrict__ scale, float *__restrict__ shift,
                                  float *__restrict__ moving_mean, float *__restrict__ moving_var, float momentum,
                                  float unbias) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float is = rsqrtf(var[c] + eps);
    const float sc = gamma[c] * is;
    invstd[c] = is;
    scale[c] = sc;
    shift[c] = mean[c];       // row 0: mean
    shift[C + c] = beta[c];   // row 1: beta   (z = (y - mean) * scale + beta)
    if (moving_mean) {
        moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * mean[c];
        moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * var[c] * unbias;
    }
}
//   backward: reduce the [blocks, C] partials (double), emit dgamma, dbeta and the apply coefficients ka, kb, kc.
//   One CTA per channel: with <= 1184 partials every thread has at most five loads, all in flight at once (this kernel
//   sits between the two passes of every batch-norm backward, 44 times per step -- its latency is on the critical path).
__global__ void __launch_bounds__(256) bn_bwd_coeffs_kernel(const float *__restrict__ part_dz, const float *__restrict__ part_dzy,
                                                            int blocks, int C, const float *__restrict__ mean,
                                                            const float *__restrict__ invstd, const float *__restrict__ gamma,
                                                            double inv_rows, int training, float *__restrict__ dgamma,
                                                            float *__restrict__ dbeta, float *__restrict__ ka,
                                                            float *__restrict__ kb, float *__restrict__ kc) {
    __shared__ double red[2][8];
    const int ch = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double s = 0.0, q = 0.0;
    int t = threadIdx.x;
    for (; t + 768 < blocks; t += 1024) {  // eight independent loads in flight per thread
        float a[4], b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[j] = part_dz[(size_t)(t + 256 * j) * C + ch]; b[j] = part_dzy[(size_t)(t + 256 * j) * C + ch]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) { s += (double)a[j]; q += (double)b[j]; }
    }
    {
        float a[3] = {0.f, 0.f, 0.f}, b[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (t + 256 * j < blocks) { a[j] = part_dz[(size_t)(t + 256 * j) * C + ch]; b[j] = part_dzy[(size_t)(t + 256 * j) * C + ch]; }
#pragma unroll
        for (int j = 0; j < 3; ++j) { s += (double)a[j]; q += (double)b[j]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (lane == 0) { red[0][wid] = s; red[1][wid] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        s = 0.0; q = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { s += red[0][w]; q += red[1][w]; }
        const double is = invstd[ch], g = gamma[ch];
        const double dzx = q * is;  // sum dz * xhat   (q = sum dz * (y - mean), already centered)
        dgamma[ch] = (float)dzx;
        dbeta[ch] = (float)s;
        if (training) {
            const double a = g * is, c = -g * is * is * (dzx * inv_rows);
            ka[ch] = (float)a;
            kc[ch] = (float)c;
            kb[ch] = (float)(-g * is * (s * inv_rows));  // applied as ka*dz + kb + kc*(y - mean)
        } else {
            ka[ch] = (float)(g * is); kb[ch] = 0.f; kc[ch] = 0.f;
        }
    }
}

static inline int ew_grid(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)kNumSMs * 16;
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

static inline bool linear_is_narrow(long long M, int K, int N) { return K <= 16 && N <= 32 && M >= 4 * NARROW_ROWS; }

template <int EPI>
static int launch_gemm(const GemmParams &p, cudaStream_t st) {
    if (EPI == EPI_STORE && linear_is_narrow(p.M, p.K, p.N)) {
        const int grid = ceil_div(p.M, NARROW_ROWS);
        if (p.K <= 8) linear_narrow_kernel<8><<<grid, 256, 0, st>>>(p);
        else linear_narrow_kernel<16><<<grid, 256, 0, st>>>(p);
    } else if (p.N <= 16) {
        dim3 grid(ceil_div(p.M, 256), ceil_div(p.N, 16));
        gemm_kernel<256, 16, 4, 4, EPI><<<grid, 256, 0, st>>>(p);
    } else {
        dim3 grid(ceil_div(p.M, 128), ceil_div(p.N, 64));
        gemm_kernel<128, 64, 8, 4, EPI><<<grid, 256, 0, st>>>(p);
    }
    PU_LAUNCH_CHECK();
    return PU_OK;
}

}  // namespace mlp
}  // namespace pu

using namespace pu;
using namespace pu::mlp;

extern "C" {

/* statistics tiling of pu_linear_fwd for a given shape: rows per tile (the kernel choice decides it) and tile count */
int pu_linear_rows_per_tile(long long M, int K, int N) {
    if (linear_is_narrow(M, K, N)) return NARROW_ROWS;
    return N <= 16 ? 256 : 128;
}
int pu_linear_row_tiles(long long M, int K, int N) { return ceil_div(M, pu_linear_rows_per_tile(M, K, N)); }

int pu_linear_fwd_ex(const float *x, int ldx, const float *w, int ldw, const float *bias, void *y, int ldy, long long M,
                     int K, int N, int accumulate, float *stat_sum, float *stat_sq, int y_dtype, pu_stream_t stream) {
    if (!x || !w || !y || M < 0 || K < 1 || N < 1 || ldx < K || ldw < N || ldy < N) return PU_ERR_INVALID_ARG;
    if ((stat_sum == nullptr) != (stat_sq == nullptr)) return PU_ERR_INVALID_ARG;
    if (M == 0) return PU_OK;
    if (ceil_div(M, 128) > 2147483647LL) return PU_ERR_UNSUPPORTED;
    if (y_dtype != PU_F32 && (y_dtype != PU_BF16 || accumulate || !linear_is_narrow(M, K, N))) return PU_ERR_UNSUPPORTED;
    GemmParams p{};
    p.A = x; p.lda = ldx; p.B = w; p.ldb = ldw; p.C = (float *)y; p.c_bf16 = y_dtype == PU_BF16; p.ldc = ldy; p.bias = bias;
    p.M = M; p.N = N; p.K = K; p.accumulate = accumulate; p.stat_sum = stat_sum; p.stat_sq = stat_sq;
    return launch_gemm<EPI_STORE>(p, (cudaStream_t)stream);
}

int pu_linear_fwd(const float *x, int ldx, const float *w, int ldw, const float *bias, float *y, int ldy, long long M,
                  int K, int N, int accumulate, float *stat_sum, float *stat_sq, pu_stream_t stream) {
    return pu_linear_fwd_ex(x, ldx, w, ldw, bias, y, ldy, M, K, N, accumulate, stat_sum, stat_sq, PU_F32, stream);
}

int pu_stats_finalize(const float *stat_sum, const float *stat_sq, int tiles, int rows_per_tile, int C, long long count,
                      float *mean, float *var, pu_stream_t stream) {
    if (!stat_sum || !stat_sq || !mean || !var || tiles < 1 || C < 1 || count < 1 || rows_per_tile < 1) return PU_ERR_INVALID_ARG;
    if ((long long)tiles != (count + rows_per_tile - 1) / rows_per_tile) return PU_ERR_INVALID_ARG;
    stats_finalize_kernel<<<C, SF_THREADS, 0, (cudaStream_t)stream>>>(stat_sum, stat_sq, tiles, C, count, rows_per_tile, mean, var,
                                                                      nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, nullptr,
                                                                      nullptr, 0.f, 1.f);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_bn_finalize_prepare(const float *stat_sum, const float *stat_sq, int tiles, int rows_per_tile, int C, long long count,
                           const float *gamma, const float *beta, float eps, float *mean, float *var, float *invstd,
                           float *scale, float *shift, float *moving_mean, float *moving_var, float momentum, float unbias,
                           pu_stream_t stream) {
    if (!stat_sum || !stat_sq || !mean || !var || !gamma || !beta || !invstd || !scale || !shift || tiles < 1 || C < 1 ||
        count < 1 || rows_per_tile < 1)
        return PU_ERR_INVALID_ARG;
    if ((long long)tiles != (count + rows_per_tile - 1) / rows_per_tile) return PU_ERR_INVALID_ARG;
    if ((moving_mean == nullptr) != (moving_var == nullptr)) return PU_ERR_INVALID_ARG;
    if (tiles <= SF_WARP_MAX_TILES)
        stats_finalize_warp_kernel<<<ceil_div(C, 8), 256, 0, (cudaStream_t)stream>>>(stat_sum, stat_sq, tiles, C, count,
                                                                                    rows_per_tile, mean, var, gamma, beta, eps,
                                                                                    invstd, scale, shift, moving_mean, moving_var,
                                                                                    momentum, unbias);
    else
        stats_finalize_kernel<<<C, SF_THREADS, 0, (cudaStream_t)stream>>>(stat_sum, stat_sq, tiles, C, count, rows_per_tile, mean,
                                                                          var, gamma, beta, eps, invstd, scale, shift, moving_mean,
                                                                          moving_var, momentum, unbias);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

static int att_args_ok(const float *x, int ldx, const float *w, long long P, int K, int d) {
    if (!x || !w || P < 0 || d < 4 || (d & 3) || ldx < d || (ldx & 3) || (((uintptr_t)x) & 15)) return PU_ERR_INVALID_ARG;
    if (K != 16) return PU_ERR_UNSUPPORTED;  // the tile layout pairs two 8-row halves per point
    return PU_OK;
}

int pu_att_pooling_fwd(const float *feature_set, int ldx, const float *w, long long P, int K, int d, float *f_agg,
                       int ldo, pu_stream_t stream) {
    int rc = att_args_ok(feature_set, ldx, w, P, K, d);
    if (rc != PU_OK) return rc;
    if (!f_agg || ldo < d || (ldo & 3) || (((uintptr_t)f_agg) & 15)) return PU_ERR_INVALID_ARG;
    if (P == 0) return PU_OK;
    GemmParams p{};
    p.A = feature_set; p.lda = ldx; p.B = w; p.ldb = d; p.M = P * K; p.N = d; p.K = d;
    p.X = feature_set; p.ldx = ldx; p.OUT = f_agg; p.ldo = ldo;
    return launch_gemm<EPI_ATT_FWD>(p, (cudaStream_t)stream);
}

int pu_att_pooling_bwd(const float *feature_set, int ldx, const float *w, const float *g_agg, int ldg, long long P,
                       int K, int d, float *d_act, int ldda, float *dx_direct, int lddx, pu_stream_t stream) {
    int rc = att_args_ok(feature_set, ldx, w, P, K, d);
    if (rc != PU_OK) return rc;
    if (!g_agg || !d_act || !dx_direct || ldg < d || ldda < d || lddx < d || ((ldg | ldda | lddx) & 3) ||
        ((((uintptr_t)g_agg) | ((uintptr_t)d_act) | ((uintptr_t)dx_direct)) & 15))
        return PU_ERR_INVALID_ARG;
    if (P == 0) return PU_OK;
    GemmParams p{};
    p.A = feature_set; p.lda = ldx; p.B = w; p.ldb = d; p.M = P * K; p.N = d; p.K = d;
    p.X = feature_set; p.ldx = ldx; p.G = g_agg; p.ldg = ldg; p.C = d_act; p.ldc = ldda; p.OUT = dx_direct; p.ldo = lddx;
    return launch_gemm<EPI_ATT_BWD>(p, (cudaStream_t)stream);
}

// thread-per-row register kernel: K <= 16 and N <= 32 natively, one of the two axes cut into slabs beyond that
// (LocSE MLPs 10 -> 64/128/256, the classifier 32 -> 4)
static inline bool wgrad_is_narrow(int K, int N) { return (K <= 16 && N <= 512) || (N <= 32 && K <= 64); }

static inline void wgrad_plan(long long M, int K, int N, int *chunks, long long *rows_per_chunk) {
    if (wgrad_is_narrow(K, N)) {
        const int slabs = ceil_div(N, 32) * ceil_div(K, 16);
        long long want = ((long long)kNumSMs * 8 + slabs - 1) / slabs;
        long long max_chunks = (M + 1023) / 1024;  // at least 1024 rows per CTA
        if (want > max_chunks) want = max_chunks;
        if (want < 1) want = 1;
        long long rpc = (M + want - 1) / want;
        *rows_per_chunk = rpc;
        *chunks = (int)((M + rpc - 1) / rpc);
        if (*chunks < 1) *chunks = 1;
        return;
    }
    const int tiles = ceil_div(K, WT) * ceil_div(N, WT);
    long long want = ((long long)kNumSMs * 8 + tiles - 1) / tiles;  // ~8 CTAs per SM in total
    long long max_chunks = (M + 255) / 256;                          // at least 256 rows per chunk
    if (want > max_chunks) want = max_chunks;
    if (want < 1) want = 1;
    long long rpc = (M + want - 1) / want;
    rpc = (rpc + WR - 1) / WR * WR;
    *rows_per_chunk = rpc;
    *chunks = (int)((M + rpc - 1) / rpc);
    if (*chunks < 1) *chunks = 1;
}

size_t pu_wgrad_workspace_bytes(long long M, int K, int N) {
    int chunks; long long rpc;
    wgrad_plan(M, K, N, &chunks, &rpc);
    return ((size_t)chunks * K * N + (size_t)chunks * N) * sizeof(float) + 256;
}

int pu_wgrad(const float *x, int ldx, const float *dy, int lddy, long long M, int K, int N, float *dw, float *db,
             int accumulate, void *workspace, size_t workspace_bytes, pu_stream_t stream) {
    if (!x || !dy || !dw || M < 0 || K < 1 || N < 1 || ldx < K || lddy < N) return PU_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        if (!accumulate) {
            PU_CUDA_TRY(cudaMemsetAsync(dw, 0, (size_t)K * N * 4, st));
            if (db) PU_CUDA_TRY(cudaMemsetAsync(db, 0, (size_t)N * 4, st));
        }
        return PU_OK;
    }
    if (!workspace || workspace_bytes < pu_wgrad_workspace_bytes(M, K, N)) return PU_ERR_WORKSPACE;
    int chunks; long long rpc;
    wgrad_plan(M, K, N, &chunks, &rpc);
    float *part = (float *)workspace;
    float *db_part = db ? part + (size_t)chunks * K * N : nullptr;
    if (wgrad_is_narrow(K, N)) {
        dim3 grid(chunks, ceil_div(N, 32), ceil_div(K, 16));
        if (K <= 8)
            wgrad_narrow_kernel<8><<<grid, 256, 0, st>>>(x, ldx, dy, lddy, M, K, N, rpc, part, db_part);
        else
            wgrad_narrow_kernel<16><<<grid, 256, 0, st>>>(x, ldx, dy, lddy, M, K, N, rpc, part, db_part);
    } else {
        dim3 grid(ceil_div(K, WT), ceil_div(N, WT), chunks);
        wgrad_kernel<<<grid, 256, 0, st>>>(x, ldx, dy, lddy, M, K, N, rpc, part, db_part);
    }
    PU_LAUNCH_CHECK();
    launch_reduce_parts(part, chunks, (long long)K * N, dw, accumulate, st);
    PU_LAUNCH_CHECK();
    if (db) {
        launch_reduce_parts(db_part, chunks, N, db, accumulate, st);
        PU_LAUNCH_CHECK();
    }
    return PU_OK;
}

int pu_bn_prepare(const float *mean, const float *var, const float *gamma, const float *beta, float eps, int C,
                  float *invstd, float *scale, float *shift, float *moving_mean, float *moving_var, float momentum,
                  float unbias, pu_stream_t stream) {
    if (!mean || !var || !gamma || !beta || !invstd || !scale || !shift || C < 1) return PU_ERR_INVALID_ARG;
    if ((moving_mean == nullptr) != (moving_var == nullptr)) return PU_ERR_INVALID_ARG;
    bn_prepare_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(mean, var, gamma, beta, eps, C, invstd, scale,
                                                                         shift, moving_mean, moving_var, momentum, unbias);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_bn_bwd_coeffs(const float *part_dz, const float *part_dzy, int blocks, int C, const float *mean,
                     const float *invstd, const float *gamma, long long rows, int training, float *dgamma, float *dbeta,
                     float *ka, float *kb, float *kc, pu_stream_t stream) {
    if (!part_dz || !part_dzy || !mean || !invstd || !gamma || !dgamma || !dbeta || !ka || !kb || !kc || blocks < 1 ||
        C < 1 || rows < 1)
        return PU_ERR_INVALID_ARG;
    bn_bwd_coeffs_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(
        part_dz, part_dzy, blocks, C, mean, invstd, gamma, 1.0 / (double)rows, training, dgamma, dbeta, ka, kb, kc);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_bn_act_fwd(const float *y, int ldy, const float *scale, const float *shift, const float *y2, int ldy2,
                  const float *scale2, const float *shift2, float slope, long long R, int C, float *out, int ldo,
                  float *out2, int ldo2, pu_stream_t stream) {
    return pu_bn_act_fwd_ex(y, ldy, PU_F32, scale, shift, y2, ldy2, PU_F32, scale2, shift2, slope, R, C, out, ldo, out2, ldo2, stream);
}

int pu_bn_act_fwd_ex(const void *y, int ldy, int y_dtype, const float *scale, const float *shift, const void *y2, int ldy2,
                     int y2_dtype, const float *scale2, const float *shift2, float slope, long long R, int C, float *out,
                     int ldo, float *out2, int ldo2, pu_stream_t stream) {
    if ((y_dtype != PU_F32 && y_dtype != PU_BF16) || (y2_dtype != PU_F32 && y2_dtype != PU_BF16)) return PU_ERR_INVALID_ARG;
    if (!y || !scale || !shift || !out || R < 0 || C < 4 || (C & 3) || ((ldy | ldo) & 3) || ldy < C || ldo < C)
        return PU_ERR_INVALID_ARG;
    if (y2 && (!scale2 || !shift2 || (ldy2 & 3) || ldy2 < C)) return PU_ERR_INVALID_ARG;
    if (out2 && ((ldo2 & 3) || ldo2 < C || (((uintptr_t)out2) & 15))) return PU_ERR_INVALID_ARG;
    if (R == 0) return PU_OK;
    if (C > 1024) return PU_ERR_UNSUPPORTED;  // one float4 column group per thread of a 256-thread CTA
    bn_act_fwd_kernel<<<ew_grid(R * (C / 4)), 256, 0, (cudaStream_t)stream>>>(y, ldy, scale, shift, y2, ldy2, scale2, shift2,
                                                                            slope, R, C, out, ldo, out2, ldo2,
                                                                            y_dtype == PU_BF16, y2_dtype == PU_BF16);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_act_bwd(const float *dout, int ldd, const float *out, int ldo, float slope, long long R, int C, float *dz, int ldz,
               pu_stream_t stream) {
    if (!dout || !out || !dz || R < 0 || C < 4 || (C & 3) || ((ldd | ldo | ldz) & 3)) return PU_ERR_INVALID_ARG;
    if (R == 0) return PU_OK;
    act_bwd_kernel<<<ew_grid(R * (C / 4)), 256, 0, (cudaStream_t)stream>>>(dout, ldd, out, ldo, slope, R, C, dz, ldz);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_bn_bwd_reduce_blocks(long long R, int C) {
    const int rpb = 256 / (C / 4) > 0 ? 256 / (C / 4) : 1;
    long long g = (R + rpb - 1) / rpb;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

int pu_bn_bwd_reduce(const float *dout, int ldd, const float *dout2, int ldd2, const float *y, int ldy, const float *scale,
                     const float *shift, float slope, long long R, int C, float *part_dz, float *part_dzy,
                     pu_stream_t stream) {
    return pu_bn_bwd_reduce_ex(dout, ldd, dout2, ldd2, y, ldy, PU_F32, scale, shift, slope, R, C, part_dz, part_dzy, stream);
}

int pu_bn_bwd_reduce_ex(const float *dout, int ldd, const float *dout2, int ldd2, const void *y, int ldy, int y_dtype,
                        const float *scale, const float *shift, float slope, long long R, int C, float *part_dz,
                        float *part_dzy, pu_stream_t stream) {
    if (y_dtype != PU_F32 && y_dtype != PU_BF16) return PU_ERR_INVALID_ARG;
    if (!dout || !y || !scale || !shift || !part_dz || !part_dzy || R < 1 || C < 4 || (C & 3) || C > 1024 * 4 ||
        ((ldd | ldy) & 3))
        return PU_ERR_INVALID_ARG;
    const int cq = C / 4;
    if (cq > 256) return PU_ERR_UNSUPPORTED;
    const int rpb = 256 / cq;
    const size_t smem = (size_t)rpb * C * 2 * sizeof(float);
    if (smem > 48 * 1024) return PU_ERR_UNSUPPORTED;
    if (dout2 && ((ldd2 & 3) || (((uintptr_t)dout2) & 15))) return PU_ERR_INVALID_ARG;
    bn_bwd_reduce_kernel<<<pu_bn_bwd_reduce_blocks(R, C), 256, smem, (cudaStream_t)stream>>>(dout, ldd, dout2, ldd2, y, ldy, scale,
                                                                                            shift, slope, R, C, part_dz, part_dzy,
                                                                                            y_dtype == PU_BF16);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

int pu_bn_bwd_apply(const float *dout, int ldd, const float *dout2, int ldd2, const float *y, int ldy, const float *scale,
                    const float *shift, float slope, const float *ka, const float *kb, const float *kc, long long R, int C,
                    float *dy, int lddy, pu_stream_t stream) {
    return pu_bn_bwd_apply_ex(dout, ldd, dout2, ldd2, y, ldy, PU_F32, scale, shift, slope, ka, kb, kc, R, C, dy, lddy, stream);
}

int pu_bn_bwd_apply_ex(const float *dout, int ldd, const float *dout2, int ldd2, const void *y, int ldy, int y_dtype,
                       const float *scale, const float *shift, float slope, const float *ka, const float *kb, const float *kc,
                       long long R, int C, float *dy, int lddy, pu_stream_t stream) {
    if (y_dtype != PU_F32 && y_dtype != PU_BF16) return PU_ERR_INVALID_ARG;
    if (!dout || !y || !scale || !shift || !ka || !kb || !kc || !dy || R < 0 || C < 4 || (C & 3) ||
        ((ldd | ldy | lddy) & 3))
        return PU_ERR_INVALID_ARG;
    if (R == 0) return PU_OK;
    if (dout2 && ((ldd2 & 3) || (((uintptr_t)dout2) & 15))) return PU_ERR_INVALID_ARG;
    if (C > 1024) return PU_ERR_UNSUPPORTED;
    bn_bwd_apply_kernel<<<ew_grid(R * (C / 4)), 256, 0, (cudaStream_t)stream>>>(dout, ldd, dout2, ldd2, y, ldy, scale, shift, slope,
                                                                              ka, kb, kc, R, C, dy, lddy, y_dtype == PU_BF16);
    PU_LAUNCH_CHECK();
    return PU_OK;
}

}  // extern "C"
