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

/* Library / build identification. */
PU_API const char *pu_version(void);
/* cudaError_t value of the last failing CUDA call made by this library on this thread (0 = none). */
PU_API int pu_last_cuda_error(void);
/* Number of kernel launches issued by this library since process start (bench.py's gpu_launches). */
PU_API unsigned long long pu_launch_count(void);

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
/* Search statistics of the last pu_knn_batch* call on this stream (device counters copied by the caller):
 * stats[0] = candidate distance evaluations, stats[1] = buckets visited, stats[2] = bucket box tests. */
PU_API int pu_knn_read_stats(const void *workspace, unsigned long long *host_stats3, pu_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POINTUNET_B200_H */
