/*
 * pointunet_b200.h -- C-ABI of libpointunet_b200.so (sm_100a).
 *
 * Drop-in boundary for the PointSegment hot path of VinAIResearch/Point-Unet.  Every entry
 * point takes plain DEVICE pointers and sizes, is asynchronous on `stream` (a cudaStream_t
 * passed as void*; NULL = legacy default stream), allocates nothing (scratch comes from the
 * caller-provided workspace), never throws or prints, and returns PU_OK or a negative
 * pu_status.  Tensors are channels-last, fp32 features / int32 indices, exactly the layouts
 * the reference's TensorFlow ops use.  "ref:" lines cite the reference interface each entry
 * replaces (paths relative to the reference repo root).
 */
#ifndef POINTUNET_B200_H
#define POINTUNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *pu_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define PU_API __attribute__((visibility("default")))
#else
#define PU_API
#endif

typedef enum {
    PU_OK = 0,
    PU_ERR_INVALID_ARG = -1, /* NULL pointer, negative size, K out of range, ... */
    PU_ERR_WORKSPACE = -2,   /* workspace NULL or smaller than pu_*_workspace_bytes() */
    PU_ERR_CUDA = -3,        /* a CUDA runtime call / launch failed (see pu_last_cuda_error) */
    PU_ERR_UNSUPPORTED = -4  /* shape outside the compiled specialisations */
} pu_status;

/* element type of a tensor that may be STORED at reduced precision (the "bf16 storage mode": the pre-normalisation
 * activations y of every 1x1 conv -- kept from the forward for the batch-norm backward, the largest saved tensors of a
 * training step -- are stored as bfloat16; arithmetic and batch statistics stay fp32).  Entry points with an `_ex` suffix
 * take it; their plain forms mean PU_F32.  Strides stay in ELEMENTS. */
typedef enum { PU_F32 = 0, PU_BF16 = 1 } pu_dtype;

/* Library / build identification. */
PU_API const char *pu_version(void);
/* cudaError_t value of the last failing CUDA call made by this library on this thread (0 = none). */
PU_API int pu_last_cuda_error(void);
/* Number of kernel launches issued by this library since process start (bench.py's gpu_launches). */
PU_API unsigned long long pu_launch_count(void);
/* CRC32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start): the checksum of TensorFlow V2 checkpoints
 * (reference: tf.train.Saver snapshots, PointSegment/RandLANet.py:101-102,180-184; restored testPancreas.py:129-132). */
PU_API unsigned int pu_crc32c(const void *data, size_t n, unsigned int crc);

/* ------------------------------------------------------------------------------------------
 * K-nearest neighbours.
 * ref: PointSegment/utils/nearest_neighbors/knn_.h:17-19  cpp_knn_batch_omp(batch_data, batch_size,
 *      npts, dim=3, queries, nqueries, K, long* batch_indices), reached through knn.pyx:71-109 and
 *      PointSegment/helper_tool.py:84-94 (DataProcessing.knn_search, which casts to int32).
 *
 *   support [B,N1,3] f32, query [B,N2,3] f32  ->  out_idx [B,N2,K] int32
 *   Row i holds the K support indices nearest to query i, ascending by
 *   (fp32 squared distance ((dx*dx)+(dy*dy))+(dz*dz) with d = q - p and no FMA, then index).
 *   If N1 < K the trailing slots are 0 (the reference leaves its zero-initialised ids).
 *   query may alias support (self-query); 1 <= K <= PU_KNN_MAX_K.
 *   Deterministic: the result does not depend on scheduling.
 * ------------------------------------------------------------------------------------------ */
#define PU_KNN_MAX_K 32
PU_API size_t pu_knn_workspace_bytes(int B, int N1, int N2, int K);
PU_API int pu_knn_batch(const float *support, const float *query, int B, int N1, int N2, int K, int32_t *out_idx,
                 void *workspace, size_t workspace_bytes, pu_stream_t stream);
/* Same search, also returning the fp32 squared distances [B,N2,K] (FLT_MAX in unfilled slots). */
PU_API int pu_knn_batch_dist(const float *support, const float *query, int B, int N1, int N2, int K, int32_t *out_idx,
                      float *out_dist, void *workspace, size_t workspace_bytes, pu_stream_t stream);
/* One pyramid level of tf_map (runPancreas.py:131-137 / runBraTS.py:147-153) from ONE search structure:
 *   out_neigh  [B,N,K] = pu_knn_batch(cloud, cloud, K)                      (neigh_idx; sub_idx is its first n_sub rows)
 *   out_interp [B,N]   = pu_knn_batch(cloud[:, :n_sub], cloud, 1)           (interp_idx), bit-identical to that call:
 * the nearest sub-cloud point of a point is the first entry of its own neighbour row with index < n_sub; the ~1 % of rows
 * without one are searched on the same structure with candidates filtered by index (boxes without a prefix point are
 * skipped).  out_unresolved (optional, uint32 [B], device): how many rows per cloud needed that search -- when the prefix is a
 * spatial region instead of a random subset most rows do, and two separate pu_knn_batch calls are the faster way
 * (both are exact; the host mirror switches on the previous call's count).  Workspace: pu_knn_workspace_bytes(B,N,N,K). */
PU_API int pu_knn_self_interp(const float *cloud, int B, int N, int K, int n_sub, int32_t *out_neigh, int32_t *out_interp,
                              unsigned *out_unresolved, void *workspace, size_t workspace_bytes, pu_stream_t stream);
/* Search statistics of the last pu_knn_batch* call on this stream (device counters copied by the caller):
 * stats[0] = candidate distance evaluations, stats[1] = buckets visited, stats[2] = bucket box tests. */
PU_API int pu_knn_read_stats(const void *workspace, unsigned long long *host_stats3, pu_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Row gathers (channels-last).  rows_per_cloud = M*K; idx is int32 [B, rows_per_cloud] with values in
 * [0, n_src); ld_* are row strides in floats (>= d) so outputs can land inside a concat buffer.
 * ref: Network.gather_neighbour  PointSegment/RandLANet.py:377-386  (pc [B,N,d], idx [B,N,K] -> [B,N,K,d])
 *      Network.nearest_interpolation  RandLANet.py:362-375          (K = 1)
 * No index validation, like the reference (TF's GPU gather does not raise on out-of-range ids).
 * ------------------------------------------------------------------------------------------ */
PU_API int pu_gather_rows_fwd(const float *src, int ld_src, int n_src, const int32_t *idx, long long rows_per_cloud,
                              int B, float *dst, int ld_dst, int d, pu_stream_t stream);
/* Inverse neighbour lists for the scatter-free backward: offsets int32 [B*n_src + 1], perm int32
 * [B*rows_per_cloud] (global row numbers, ascending inside each segment -> deterministic sums).  Built as a counting
 * sort; the workspace holds the unsorted lists, B*n_src + 1 counters and the scan scratch: pu_inverse_workspace_bytes
 * covers n_src <= 4 * rows_per_cloud, callers with more targets than that add 4 * (B*n_src + 1) bytes. */
PU_API size_t pu_inverse_workspace_bytes(int B, long long rows_per_cloud);
PU_API int pu_build_inverse(const int32_t *idx, long long rows_per_cloud, int B, int n_src, int32_t *offsets,
                            int32_t *perm, void *workspace, size_t workspace_bytes, pu_stream_t stream);
/* grad_src[j,:] (+)= sum over the inverse list of j of grad_out[row,:]   (gradient of any row gather) */
PU_API int pu_segment_sum(const float *grad_out, int ld_go, const int32_t *offsets, const int32_t *perm,
                          long long n_targets, float *grad_src, int ld_gs, int d, int accumulate, pu_stream_t stream);

/* ref: Network.relative_pos_encoding  RandLANet.py:337-343
 *   xyz [B,N,3], idx [B,N,K] -> out [B,N,K,10] = [dist, rel(3), tile(3), neighbour(3)]; no gradient (xyz is data). */
PU_API int pu_relative_pos_encoding_fwd(const float *xyz, const int32_t *idx, int B, int N, int K, float *out,
                                        pu_stream_t stream);

/* ref: building_block, position branch  RandLANet.py:323-326:  relative_pos_encoding -> conv2d(10 -> h, 'mlp1') ->
 *   BN(0.99, 1e-6) -> LeakyReLU(0.2), as recompute kernels (csrc/locse_mlp.cu): neither the 10-channel LocSE rows nor the
 *   pre-normalisation tensor nor its gradient are ever stored.  h = power of two in [4, 1024].
 *   The kernels read the cloud as padded points xyz4 [B,N,4] = (x, y, z, 0), 16-byte aligned (one 128-bit access per end
 *   point instead of three scalar gathers): pu_locse_pack_xyz makes that copy from xyz [n_points, 3].
 *   pu_locse_moments:    mom[0:10] = column sums of the LocSE rows over all B*N*K rows, mom[10:65] = centred second-moment
 *                        sums (upper triangle, row-major); the batch statistics of the conv output follow from them.
 *   pu_locse_bn_prepare: coef (5h + 110 floats, 16-byte aligned): scale | t | invstd | mean_y | var_y | xbar(10) | Cov(100).
 *                        training != 0: batch statistics (moving_mean / moving_var, when given, receive the momentum update
 *                        with `unbias`, like pu_bn_finalize_prepare); training == 0: the moving statistics are used.
 *   pu_locse_mlp_fwd:    out[row, 0:h] (row stride ldo; optionally also out2) = lrelu(scale * ((x - xbar) W) + t).
 *   pu_locse_mlp_bwd:    backward given dz (+ dz2, summed on the fly) and the coef of the forward (same `training`): dW [10,h]
 *                        ((+)= when accumulate_dw), dbias (optional; identically 0 in training mode), dgamma, dbeta [h]; there
 *                        is no dgrad (xyz is data).  Deterministic. */
PU_API int pu_locse_mlp_supported(int K, int h);
PU_API size_t pu_locse_mlp_workspace_bytes(int h);
PU_API int pu_locse_pack_xyz(const float *xyz, long long n_points, float *xyz4, pu_stream_t stream);
PU_API int pu_locse_moments(const float *xyz4, const int32_t *idx, int B, int N, int K, float *mom, void *workspace,
                            size_t workspace_bytes, pu_stream_t stream);
PU_API int pu_locse_bn_prepare(const float *mom, long long count, const float *w, int h, const float *bias,
                               const float *gamma, const float *beta, float eps, int training, float *moving_mean,
                               float *moving_var, float momentum, float unbias, float *coef, pu_stream_t stream);
PU_API int pu_locse_mlp_fwd(const float *xyz4, const int32_t *idx, int B, int N, int K, const float *w, int h,
                            const float *coef, float slope, float *out, int ldo, float *out2, int ldo2, pu_stream_t stream);
PU_API int pu_locse_mlp_bwd(const float *xyz4, const int32_t *idx, int B, int N, int K, const float *w, int h,
                            const float *coef, const float *gamma, const float *bias, int training, float slope,
                            const float *dz, int ldz, const float *dz2, int ldz2, float *dw, int accumulate_dw, float *dbias, float *dgamma, float *dbeta,
                            void *workspace, size_t workspace_bytes, pu_stream_t stream);

/* ref: Network.random_sample  RandLANet.py:345-360   out[b,m,:] = max_k feat[b, pool_idx[b,m,k], :]
 *   ties (uint8 [B*M, d], optional) = how many neighbours attain the max; the backward splits the gradient
 *   evenly among exact ties like tf.reduce_max and walks the inverse list of pool_idx (scatter-free). */
PU_API int pu_random_sample_fwd(const float *feat, int ld_f, int n_src, const int32_t *pool_idx, int B, int M, int K,
                                float *out, int ld_o, unsigned char *ties, int d, pu_stream_t stream);
PU_API int pu_random_sample_bwd(const float *feat, int ld_f, const float *out, int ld_o, const unsigned char *ties,
                                const float *g_out, int ld_g, const int32_t *offsets, const int32_t *perm,
                                long long n_targets, int K, float *g_feat, int ld_gf, int d, pu_stream_t stream);

/* ref: point2prod  PointSegment/testPancreas.py:71-85, testBraTS.py:83-101 (test-mode fusion of per-point
 *   probabilities into a dense volume).  probs [n,C] f32, xyz_origin int32 [*,3] = (x,y,z) voxel coordinates,
 *   point_idx (optional int32 [n]: row of xyz_origin for point i -- the BraTS `test_probs[p_idx] = probs` step,
 *   testBraTS.py:226-231) -> volume f32 [Z,Y,X,C], i.e. `volume[z][x][y] = prob[i]` followed by
 *   np.moveaxis(volume, 1, 2); untouched voxels are 0; if several points hit one voxel the last one wins, as in
 *   the reference's sequential loop.  Out-of-range coordinates are skipped. */
PU_API size_t pu_point2prod_workspace_bytes(int Z, int X, int Y);
PU_API int pu_point2prod(const float *probs, const int32_t *xyz_origin, const int32_t *point_idx, int n, int C, int Z,
                         int X, int Y, float *volume, void *workspace, size_t workspace_bytes, pu_stream_t stream);

/* ref: genSegmentation  utils/genSegmentationPancreas.py:67-77, utils/genSegmentationBraTS.py:67-78: the label volume
 *   seg = np.argmax(prob_volume, axis=-1).astype(uint8) (first maximum wins; voxels no point fell into are 0), with the
 *   BraTS relabelling 3 -> 4 as (remap_from, remap_to) = (3, 4) (pass (-1, 0) for none).
 *   pu_point2label fuses point2prod + argmax: labels u8 [Z,Y,X] straight from the per-point probabilities, the dense
 *   fp32 probability volume is never written (same workspace as pu_point2prod).  pu_volume_argmax is the stand-alone
 *   argmax over the last axis of an existing volume [nvox, C]. */
PU_API int pu_point2label(const float *probs, const int32_t *xyz_origin, const int32_t *point_idx, int n, int C, int Z,
                          int X, int Y, int remap_from, int remap_to, unsigned char *labels, void *workspace,
                          size_t workspace_bytes, pu_stream_t stream);
PU_API int pu_volume_argmax(const float *volume, long long nvox, int C, int remap_from, int remap_to,
                            unsigned char *labels, pu_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Shared MLP = 1x1 convolution over channels-last rows.
 * ref: helper_tf_util.conv2d  PointSegment/helper_tf_util.py:115-170 ; conv2d_transpose :173-250 ;
 *      tf.layers.dense RandLANet.py:114.   y[M,N] (+)= x[M,K] w[K,N] + bias.
 *   stat_sum/stat_sq (optional, [pu_linear_row_tiles(M,K,N), N]) receive per-tile (sum, centred M2) of every column of y
 *   (tiles of pu_linear_rows_per_tile(M,K,N) rows; 128 rows for the tensor-core path);
 *   pu_stats_finalize turns them into the batch-norm mean and BIASED variance (training mode, :166). */
PU_API int pu_linear_rows_per_tile(long long M, int K, int N);
PU_API int pu_linear_row_tiles(long long M, int K, int N);
PU_API int pu_linear_fwd(const float *x, int ldx, const float *w, int ldw, const float *bias, float *y, int ldy,
                         long long M, int K, int N, int accumulate, float *stat_sum, float *stat_sq,
                         pu_stream_t stream);
PU_API int pu_linear_fwd_ex(const float *x, int ldx, const float *w, int ldw, const float *bias, void *y, int ldy,
                            long long M, int K, int N, int accumulate, float *stat_sum, float *stat_sq, int y_dtype,
                            pu_stream_t stream);   /* PU_BF16: narrow shapes only (K <= 16, N <= 32), no accumulate */
PU_API int pu_stats_finalize(const float *stat_sum, const float *stat_sq, int tiles, int rows_per_tile, int C,
                             long long count, float *mean, float *var, pu_stream_t stream);
/* Tensor-core (tcgen05 + TMEM) form of pu_linear_fwd for K >= 32, N >= 32: y[M,N] (+)= x[M,K] wt[N,K]^T + bias, where
 * wt is the TRANSPOSED weight (K-major).  mode 3 = 3xTF32 (hi/lo split, fp32-class accuracy -- the parity path),
 * mode 1 = plain TF32.  stat_sum/stat_m2: per-128-row-tile batch-norm partials as in pu_linear_fwd.
 * *error_flag (device int, optional) is set to 1 if an internal barrier wait timed out. */
PU_API int pu_tc_linear_supported(long long M, int K, int N, int ldx, int ldwt, int ldy);
PU_API size_t pu_tc_workspace_bytes(int K, int N); /* scratch for the packed weight image (streamed-weight kernels) */
PU_API int pu_tc_linear_fwd(const float *x, int ldx, const float *wt, int ldwt, const float *bias, float *y, int ldy,
                            long long M, int K, int N, int accumulate, float *stat_sum, float *stat_m2, int mode,
                            int *error_flag, void *workspace, size_t workspace_bytes, pu_stream_t stream);
PU_API int pu_tc_linear_fwd_ex(const float *x, int ldx, const float *wt, int ldwt, const float *bias, void *y, int ldy,
                               long long M, int K, int N, int accumulate, float *stat_sum, float *stat_m2, int mode,
                               int *error_flag, void *workspace, size_t workspace_bytes, int y_dtype,
                               pu_stream_t stream);   /* PU_BF16: y stored as bf16 (no accumulate); statistics from the fp32 values */
/* Tensor-core forms of pu_att_pooling_fwd / _bwd (channel width d >= 32, K = 16): same contract, but `wt` is the
 * TRANSPOSED FC kernel [d_out, d_in] (K-major); the softmax-over-K epilogue runs out of TMEM. */
PU_API int pu_tc_att_supported(int K, int d, int ldx);
PU_API int pu_tc_att_pooling_fwd(const float *feature_set, int ldx, const float *wt, long long P, int K, int d,
                                 float *f_agg, int ldo, int mode, int *error_flag, void *workspace,
                                 size_t workspace_bytes, pu_stream_t stream);
PU_API int pu_tc_att_pooling_bwd(const float *feature_set, int ldx, const float *wt, const float *g_agg, int ldg,
                                 long long P, int K, int d, float *d_act, int ldda, float *dx_direct, int lddx,
                                 int mode, int *error_flag, void *workspace, size_t workspace_bytes,
                                 pu_stream_t stream);
/* d = 64: the dgrad through the FC fused into the same kernel -- `dx` receives the complete gradient g s + d_act w^T, the
 * separate accumulate GEMM (pu_tc_linear_fwd with accumulate) is not needed; `w` = FC kernel [d,d], `wt` = its transpose. */
PU_API int pu_tc_att_bwd_fused_supported(int K, int d, int ldx);
PU_API int pu_tc_att_pooling_bwd_fused(const float *feature_set, int ldx, const float *wt, const float *w, const float *g_agg,
                                       int ldg, long long P, int K, int d, float *d_act, int ldda, float *dx, int lddx,
                                       int mode, int *error_flag, pu_stream_t stream);
/* Tensor-core weight gradient (tcgen05, MN-major operands, all rows of a CTA accumulated in TMEM): same contract as
 * pu_wgrad for Kin, N >= 32 and M >= 4096; db is produced through a ones row and needs Kin % 128 != 0. */
PU_API int pu_tc_wgrad_supported(long long M, int Kin, int N, int ldx, int lddy, int want_db);
PU_API size_t pu_tc_wgrad_workspace_bytes(long long M, int Kin, int N);
PU_API int pu_tc_wgrad(const float *x, int ldx, const float *dy, int lddy, long long M, int Kin, int N, float *dw,
                       float *db, int accumulate, int mode, void *workspace, size_t workspace_bytes, int *error_flag,
                       pu_stream_t stream);
/* dw[K,N] (+)= x^T dy ; db[N] (+)= column sums of dy (db may be NULL).  Deterministic. */
PU_API size_t pu_wgrad_workspace_bytes(long long M, int K, int N);
PU_API int pu_wgrad(const float *x, int ldx, const float *dy, int lddy, long long M, int K, int N, float *dw,
                    float *db, int accumulate, void *workspace, size_t workspace_bytes, pu_stream_t stream);
/* out = leaky_relu_slope(y*scale + shift [+ y2*scale2 + shift2]) per channel (slope 1 = no activation);
 * the two-input form is the residual sum of dilated_res_block (RandLANet.py:317-321). */
PU_API int pu_bn_act_fwd(const float *y, int ldy, const float *scale, const float *shift, const float *y2, int ldy2,
                         const float *scale2, const float *shift2, float slope, long long R, int C, float *out,
                         int ldo, float *out2, int ldo2, pu_stream_t stream);
PU_API int pu_bn_act_fwd_ex(const void *y, int ldy, int y_dtype, const float *scale, const float *shift, const void *y2,
                            int ldy2, int y2_dtype, const float *scale2, const float *shift2, float slope, long long R,
                            int C, float *out, int ldo, float *out2, int ldo2, pu_stream_t stream);
/* (`shift` is a [2,C] array: row 0 = mean, row 1 = beta -- z = (y - mean)*scale + beta.  `out2` optionally receives a
 *  second copy of the result, e.g. the right half of an LFA concat buffer, RandLANet.py:328,333.) */
/* Per-channel batch-norm coefficients, one launch each (tf.layers.batch_normalization, momentum 0.99, eps 1e-6):
 *   forward : invstd = rsqrt(var+eps), scale = gamma*invstd, shift = beta - mean*scale, and (if moving_* given)
 *             moving = momentum*moving + (1-momentum)*batch  (variance times `unbias`, n/(n-1) for TF's fused path);
 *   backward: reduces the pu_bn_bwd_reduce partials and emits dgamma, dbeta and ka/kb/kc for pu_bn_bwd_apply. */
PU_API int pu_bn_prepare(const float *mean, const float *var, const float *gamma, const float *beta, float eps, int C,
                         float *invstd, float *scale, float *shift, float *moving_mean, float *moving_var,
                         float momentum, float unbias, pu_stream_t stream);
/* pu_stats_finalize + pu_bn_prepare in ONE launch (training mode: the statistics come straight from the per-tile partials of
 * the producing linear kernel; 44 batch norms per step, each saved launch sits on the critical path). */
PU_API int pu_bn_finalize_prepare(const float *stat_sum, const float *stat_sq, int tiles, int rows_per_tile, int C,
                                  long long count, const float *gamma, const float *beta, float eps, float *mean,
                                  float *var, float *invstd, float *scale, float *shift, float *moving_mean,
                                  float *moving_var, float momentum, float unbias, pu_stream_t stream);
PU_API int pu_bn_bwd_coeffs(const float *part_dz, const float *part_dzy, int blocks, int C, const float *mean,
                            const float *invstd, const float *gamma, long long rows, int training, float *dgamma,
                            float *dbeta, float *ka, float *kb, float *kc, pu_stream_t stream);
/* dz = dout * (out > 0 ? 1 : slope) */
PU_API int pu_act_bwd(const float *dout, int ldd, const float *out, int ldo, float slope, long long R, int C,
                      float *dz, int ldz, pu_stream_t stream);
/* batch-norm backward: pass 1 reduces sum(dz) and sum(dz*y) per channel into [pu_bn_bwd_reduce_blocks, C]
 * partials (dz = dout * lrelu'(y*scale+shift)); pass 2 applies dy = ka*dz + kb + kc*y. */
PU_API int pu_bn_bwd_reduce_blocks(long long R, int C);
/* dout2 (optional) is a second upstream gradient added on the fly: dz = (dout + dout2) * lrelu'(...) */
PU_API int pu_bn_bwd_reduce(const float *dout, int ldd, const float *dout2, int ldd2, const float *y, int ldy,
                            const float *scale, const float *shift, float slope, long long R, int C, float *part_dz,
                            float *part_dzy, pu_stream_t stream);
PU_API int pu_bn_bwd_apply(const float *dout, int ldd, const float *dout2, int ldd2, const float *y, int ldy,
                           const float *scale, const float *shift, float slope, const float *ka, const float *kb,
                           const float *kc, long long R, int C, float *dy, int lddy, pu_stream_t stream);
PU_API int pu_bn_bwd_reduce_ex(const float *dout, int ldd, const float *dout2, int ldd2, const void *y, int ldy, int y_dtype,
                               const float *scale, const float *shift, float slope, long long R, int C, float *part_dz,
                               float *part_dzy, pu_stream_t stream);
PU_API int pu_bn_bwd_apply_ex(const float *dout, int ldd, const float *dout2, int ldd2, const void *y, int ldy, int y_dtype,
                              const float *scale, const float *shift, float slope, const float *ka, const float *kb,
                              const float *kc, long long R, int C, float *dy, int lddy, pu_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ref: Network.att_pooling  RandLANet.py:388-401 (up to f_agg; the trailing conv2d is pu_linear_fwd)
 *   feature_set [P, K, d] (P = B*N points, row stride ldx), w [d,d] (tf.layers.dense kernel, no bias)
 *   f_agg[p,c] = sum_k x[p,k,c] * softmax_k(x[p,k,:] w)[c]          -- one fused kernel, K must be 16.
 * Backward: d_act [P*K, d] = s*(g*x - sum_k g*x*s), dx_direct [P*K, d] = g*s; the caller finishes with
 *   dx = dx_direct + d_act w^T  (pu_linear_fwd, accumulate)  and  dw = x^T d_act  (pu_wgrad).
 * ------------------------------------------------------------------------------------------ */
PU_API int pu_att_pooling_fwd(const float *feature_set, int ldx, const float *w, long long P, int K, int d,
                              float *f_agg, int ldo, pu_stream_t stream);
PU_API int pu_att_pooling_bwd(const float *feature_set, int ldx, const float *w, const float *g_agg, int ldg,
                              long long P, int K, int d, float *d_act, int ldda, float *dx_direct, int lddx,
                              pu_stream_t stream);

/* The same op for the 16-channel level (K = 16, d = 16: Encoder_layer_0, RandLANet.py:323-335 with d_out = 16), one
 * kernel each way.  pu_att16_bwd is the COMPLETE gradient in one pass over x: dx [P*16, 16] (row stride lddx) and
 * dw [16,16] (added to its content when `accumulate`), no d_act / g*s intermediates in HBM.  `w` is the [16,16] kernel in
 * device memory; each call copies it (stream-ordered) into one of four constant-memory slots used round robin, so at most
 * four calls may be in flight on DIFFERENT streams at once.  workspace: per-CTA partials of dw. */
PU_API int pu_att16_supported(int K, int d, int ldx);
PU_API size_t pu_att16_workspace_bytes(long long P);
PU_API int pu_att16_fwd(const float *feature_set, int ldx, const float *w, long long P, float *f_agg, int ldo,
                        pu_stream_t stream);
PU_API int pu_att16_bwd(const float *feature_set, int ldx, const float *w, const float *g_agg, int ldg, long long P,
                        float *dx, int lddx, float *dw, int accumulate, void *workspace, size_t workspace_bytes,
                        pu_stream_t stream);
/* The same kernels with the 16 channels of a row in TWO tensors (x_lo: channels 0-7, x_hi: channels 8-15; likewise dx): the
 * two halves of building_block's concat (RandLANet.py:328,333) kept as separate contiguous tensors -- at 8 channels a half
 * row is 32 bytes, and reading / writing it inside a 64-byte row costs a full DRAM burst per half. */
PU_API int pu_att16_supported_split(int K, int d, int ld_lo, int ld_hi);
PU_API int pu_att16_fwd_split(const float *x_lo, int ld_lo, const float *x_hi, int ld_hi, const float *w, long long P,
                              float *f_agg, int ldo, pu_stream_t stream);
PU_API int pu_att16_bwd_split(const float *x_lo, int ld_lo, const float *x_hi, int ld_hi, const float *w, const float *g_agg,
                              int ldg, long long P, float *dx_lo, int lddx_lo, float *dx_hi, int lddx_hi, float *dw,
                              int accumulate, void *workspace, size_t workspace_bytes, pu_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POINTUNET_B200_H */
