/*
 * oracle/knn_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * CPU restatement, in plain C, of the K-nearest-neighbour search the reference runs on the
 * host: a nanoflann kd-tree (v0x123) built per cloud and queried point by point.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's shared object.
 *
 * Parity pin: this restatement is checked bit-for-bit (indices, every row, ties included)
 * against the reference's own C++ compiled from /root/reference into oracle/_ref/
 * (tests/test_oracle_knn.py) and against the committed fixtures in tests/golden/.
 *
 * What follows what (paths relative to /root/reference/PointSegment/utils/nearest_neighbors):
 *   l2_eval            nanoflann.hpp:323-348   L2_Adaptor::evalMetric, dim=3 tail loop
 *   rs_add             nanoflann.hpp:114-141   KNNResultSet::addPoint (strict '>' shift)
 *   min_max            nanoflann.hpp:897-908   computeMinMax
 *   plane_split        nanoflann.hpp:1016-1043 planeSplit (3-way partition)
 *   middle_split       nanoflann.hpp:966-1005  middleSplit_
 *   divide_tree        nanoflann.hpp:916-964   divideTree (leaf_max_size 10)
 *   build_index        nanoflann.hpp:1216-1226 buildIndex + computeBoundingBox :1321-1343
 *   search_level       nanoflann.hpp:1350-1408 searchLevel (eps = 0)
 *   find_neighbors     nanoflann.hpp:1243-1258 findNeighbors + computeInitialDistances :1045-1061
 *   pu_oracle_knn_batch  knn_.cxx:104-135      cpp_knn_batch_omp (OpenMP over clouds only)
 *
 * Two tie rules:
 *   tie_rule 0  "nanoflann": first visited wins among equal distances (kd traversal order).
 *   tie_rule 1  "canonical": order by (distance, index) ascending; this is the rule the
 *               B200 path states and implements.  The traversal is the same, the leaf test
 *               becomes '<=' and insertion compares (dist, index) lexicographically; the
 *               branch-pruning bound gets a 1e-5 relative slack so that fp32 rounding of the
 *               incrementally maintained bound can never drop a boundary tie.
 *
 * Distance arithmetic is fp32, d = q - p, ((dx*dx)+(dy*dy))+(dz*dz), every operation rounded
 * separately (compile with -ffp-contract=off, no -march=native, no -ffast-math).
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LEAF_MAX 10 /* KDTreeTableAdaptor ctor default used by knn_.cxx:115 */

typedef struct {
    int child1, child2; /* -1/-1 marks a leaf */
    int left, right;    /* leaf: range in vind */
    int divfeat;
    float divlow, divhigh;
} node_t;

typedef struct {
    const float *pts; /* [n,3] */
    int n;
    int *vind;
    node_t *nodes;
    int n_nodes, cap_nodes;
    float root_lo[3], root_hi[3];
} tree_t;

typedef struct {
    int64_t *ids;
    float *dists;
    int cap, count;
    int tie_rule;
    long evals; /* distance evaluations, for the bench's evals/s figure */
} rs_t;

static inline float pt(const tree_t *t, int i, int c) { return t->pts[(size_t)i * 3 + c]; }

static inline float l2_eval(const float *q, const float *p)
{
    float r = 0.0f;
    for (int c = 0; c < 3; ++c) {
        const float diff = q[c] - p[c];
        r += diff * diff;
    }
    return r;
}

static inline void rs_init(rs_t *rs)
{
    rs->count = 0;
    if (rs->cap) rs->dists[rs->cap - 1] = FLT_MAX;
}

static inline float rs_worst(const rs_t *rs) { return rs->dists[rs->cap - 1]; }

static inline void rs_add(rs_t *rs, float dist, int64_t index)
{
    int i;
    for (i = rs->count; i > 0; --i) {
        int shift;
        if (rs->tie_rule == 0)
            shift = rs->dists[i - 1] > dist;
        else
            shift = (rs->dists[i - 1] > dist) || (rs->dists[i - 1] == dist && rs->ids[i - 1] > index);
        if (!shift) break;
        if (i < rs->cap) {
            rs->dists[i] = rs->dists[i - 1];
            rs->ids[i] = rs->ids[i - 1];
        }
    }
    if (i < rs->cap) {
        rs->dists[i] = dist;
        rs->ids[i] = index;
    }
    if (rs->count < rs->cap) rs->count++;
}

static void min_max(const tree_t *t, const int *ind, int count, int c, float *mn, float *mx)
{
    *mn = *mx = pt(t, ind[0], c);
    for (int i = 1; i < count; ++i) {
        const float v = pt(t, ind[i], c);
        if (v < *mn) *mn = v;
        if (v > *mx) *mx = v;
    }
}

static void plane_split(const tree_t *t, int *ind, int count, int cutfeat, float cutval, int *lim1, int *lim2)
{
    /* nanoflann uses unsigned IndexType; '!right' guards the wrap at zero. Signed ints here,
     * the guards are kept so the control flow is identical. */
    long left = 0, right = (long)count - 1;
    for (;;) {
        while (left <= right && pt(t, ind[left], cutfeat) < cutval) ++left;
        while (right && left <= right && pt(t, ind[right], cutfeat) >= cutval) --right;
        if (left > right || !right) break;
        int tmp = ind[left]; ind[left] = ind[right]; ind[right] = tmp;
        ++left; --right;
    }
    *lim1 = (int)left;
    right = (long)count - 1;
    for (;;) {
        while (left <= right && pt(t, ind[left], cutfeat) <= cutval) ++left;
        while (right && left <= right && pt(t, ind[right], cutfeat) > cutval) --right;
        if (left > right || !right) break;
        int tmp = ind[left]; ind[left] = ind[right]; ind[right] = tmp;
        ++left; --right;
    }
    *lim2 = (int)left;
}

static void middle_split(const tree_t *t, int *ind, int count, int *index, int *cutfeat, float *cutval,
                         const float lo[3], const float hi[3])
{
    const float EPS = 0.00001f;
    float max_span = hi[0] - lo[0];
    for (int i = 1; i < 3; ++i) {
        const float span = hi[i] - lo[i];
        if (span > max_span) max_span = span;
    }
    float max_spread = -1.0f;
    *cutfeat = 0;
    for (int i = 0; i < 3; ++i) {
        const float span = hi[i] - lo[i];
        if (span > (1 - EPS) * max_span) {
            float mn, mx;
            min_max(t, ind, count, i, &mn, &mx);
            const float spread = mx - mn;
            if (spread > max_spread) {
                *cutfeat = i;
                max_spread = spread;
            }
        }
    }
    const float split_val = (lo[*cutfeat] + hi[*cutfeat]) / 2;
    float mn, mx;
    min_max(t, ind, count, *cutfeat, &mn, &mx);
    if (split_val < mn) *cutval = mn;
    else if (split_val > mx) *cutval = mx;
    else *cutval = split_val;

    int lim1, lim2;
    plane_split(t, ind, count, *cutfeat, *cutval, &lim1, &lim2);
    if (lim1 > count / 2) *index = lim1;
    else if (lim2 < count / 2) *index = lim2;
    else *index = count / 2;
}

static int new_node(tree_t *t)
{
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
        t->nodes = (node_t *)realloc(t->nodes, (size_t)t->cap_nodes * sizeof(node_t));
    }
    return t->n_nodes++;
}

static int divide_tree(tree_t *t, int left, int right, float lo[3], float hi[3])
{
    const int id = new_node(t);
    if ((right - left) <= LEAF_MAX) {
        t->nodes[id].child1 = t->nodes[id].child2 = -1;
        t->nodes[id].left = left;
        t->nodes[id].right = right;
        for (int i = 0; i < 3; ++i) lo[i] = hi[i] = pt(t, t->vind[left], i);
        for (int k = left + 1; k < right; ++k)
            for (int i = 0; i < 3; ++i) {
                const float v = pt(t, t->vind[k], i);
                if (lo[i] > v) lo[i] = v;
                if (hi[i] < v) hi[i] = v;
            }
    } else {
        int idx, cutfeat;
        float cutval;
        middle_split(t, t->vind + left, right - left, &idx, &cutfeat, &cutval, lo, hi);
        float llo[3], lhi[3], rlo[3], rhi[3];
        memcpy(llo, lo, sizeof(llo)); memcpy(lhi, hi, sizeof(lhi));
        memcpy(rlo, lo, sizeof(rlo)); memcpy(rhi, hi, sizeof(rhi));
        lhi[cutfeat] = cutval;
        const int c1 = divide_tree(t, left, left + idx, llo, lhi);
        rlo[cutfeat] = cutval;
        const int c2 = divide_tree(t, left + idx, right, rlo, rhi);
        node_t *nd = &t->nodes[id]; /* re-fetch: nodes may have been reallocated */
        nd->child1 = c1;
        nd->child2 = c2;
        nd->divfeat = cutfeat;
        nd->divlow = lhi[cutfeat];
        nd->divhigh = rlo[cutfeat];
        for (int i = 0; i < 3; ++i) {
            lo[i] = llo[i] < rlo[i] ? llo[i] : rlo[i];
            hi[i] = lhi[i] > rhi[i] ? lhi[i] : rhi[i];
        }
    }
    return id;
}

static void build_index(tree_t *t, const float *pts, int n)
{
    memset(t, 0, sizeof(*t));
    t->pts = pts;
    t->n = n;
    t->vind = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    for (int i = 0; i < n; ++i) t->vind[i] = i;
    if (n == 0) return;
    for (int i = 0; i < 3; ++i) t->root_lo[i] = t->root_hi[i] = pt(t, 0, i);
    for (int k = 1; k < n; ++k)
        for (int i = 0; i < 3; ++i) {
            const float v = pt(t, k, i);
            if (v < t->root_lo[i]) t->root_lo[i] = v;
            if (v > t->root_hi[i]) t->root_hi[i] = v;
        }
    divide_tree(t, 0, n, t->root_lo, t->root_hi);
}

static void free_index(tree_t *t)
{
    free(t->vind);
    free(t->nodes);
}

static void search_level(const tree_t *t, rs_t *rs, const float *q, int node, float mindistsq, float dists[3])
{
    const node_t *nd = &t->nodes[node];
    if (nd->child1 < 0 && nd->child2 < 0) {
        const float worst = rs_worst(rs); /* sampled once at leaf entry, nanoflann.hpp:1357 */
        for (int i = nd->left; i < nd->right; ++i) {
            const int index = t->vind[i];
            const float dist = l2_eval(q, t->pts + (size_t)index * 3);
            rs->evals++;
            if (rs->tie_rule == 0 ? (dist < worst) : (dist <= worst)) rs_add(rs, dist, index);
        }
        return;
    }
    const int idx = nd->divfeat;
    const float val = q[idx];
    const float diff1 = val - nd->divlow;
    const float diff2 = val - nd->divhigh;
    int best, other;
    float cut_dist;
    if ((diff1 + diff2) < 0) {
        best = nd->child1; other = nd->child2;
        cut_dist = (val - nd->divhigh) * (val - nd->divhigh);
    } else {
        best = nd->child2; other = nd->child1;
        cut_dist = (val - nd->divlow) * (val - nd->divlow);
    }
    search_level(t, rs, q, best, mindistsq, dists);
    const float dst = dists[idx];
    mindistsq = mindistsq + cut_dist - dst;
    dists[idx] = cut_dist;
    const float bound = rs->tie_rule == 0 ? mindistsq * 1.0f : mindistsq * (1.0f - 1e-5f);
    if (bound <= rs_worst(rs)) search_level(t, rs, q, other, mindistsq, dists);
    dists[idx] = dst;
}

static void find_neighbors(const tree_t *t, rs_t *rs, const float *q)
{
    if (t->n == 0) return;
    float dists[3] = {0, 0, 0};
    float distsq = 0.0f;
    for (int i = 0; i < 3; ++i) {
        if (q[i] < t->root_lo[i]) { dists[i] = (q[i] - t->root_lo[i]) * (q[i] - t->root_lo[i]); distsq += dists[i]; }
        if (q[i] > t->root_hi[i]) { dists[i] = (q[i] - t->root_hi[i]) * (q[i] - t->root_hi[i]); distsq += dists[i]; }
    }
    search_level(t, rs, q, 0, distsq, dists);
}

/*
 * Batched K-NN, restating cpp_knn_batch_omp (knn_.cxx:104-135).
 *   support [B,N1,3] f32, query [B,N2,3] f32 -> out_idx int64 [B,N2,K] (must be zero-initialised by the
 *   caller like knn.pyx:93; slots beyond N1 stay untouched), out_dist f32 [B,N2,K] or NULL.
 * Returns the total number of distance evaluations.
 */
long pu_oracle_knn_batch(const float *support, long B, long N1, const float *query, long N2, long K,
                         int64_t *out_idx, float *out_dist, int tie_rule, int use_omp)
{
    long total_evals = 0;
#pragma omp parallel for reduction(+ : total_evals) if (use_omp)
    for (long b = 0; b < B; ++b) {
        tree_t t;
        build_index(&t, support + (size_t)b * N1 * 3, (int)N1);
        int64_t *ids = (int64_t *)calloc((size_t)K, sizeof(int64_t));
        float *ds = (float *)calloc((size_t)K, sizeof(float));
        rs_t rs;
        rs.ids = ids; rs.dists = ds; rs.cap = (int)K; rs.tie_rule = tie_rule; rs.evals = 0;
        for (long i = 0; i < N2; ++i) {
            rs_init(&rs);
            find_neighbors(&t, &rs, query + ((size_t)b * N2 + i) * 3);
            int64_t *o = out_idx + ((size_t)b * N2 + i) * K;
            /* the reference copies all K slots of a buffer that persists across queries and starts
             * zeroed (knn_.cxx:120-130); slots >= N1 are therefore always zero. */
            for (long j = 0; j < K; ++j) o[j] = j < rs.count ? ids[j] : 0;
            if (out_dist) {
                float *od = out_dist + ((size_t)b * N2 + i) * K;
                for (long j = 0; j < K; ++j) od[j] = j < rs.count ? ds[j] : FLT_MAX;
            }
        }
        total_evals += rs.evals;
        free(ids); free(ds);
        free_index(&t);
    }
    return total_evals;
}

/*
 * Exhaustive (distance, index)-ordered K-NN: the definition of the canonical tie rule, O(N1*N2).
 * Used to pin the kd-tree canonical mode at small sizes.
 */
void pu_oracle_knn_brute(const float *support, long B, long N1, const float *query, long N2, long K,
                         int64_t *out_idx, float *out_dist)
{
#pragma omp parallel for collapse(2)
    for (long b = 0; b < B; ++b)
        for (long i = 0; i < N2; ++i) {
            int64_t ids[64];
            float ds[64];
            rs_t rs;
            rs.ids = ids; rs.dists = ds; rs.cap = (int)K; rs.tie_rule = 1; rs.evals = 0;
            rs_init(&rs);
            const float *q = query + ((size_t)b * N2 + i) * 3;
            for (long j = 0; j < N1; ++j) {
                const float dist = l2_eval(q, support + ((size_t)b * N1 + j) * 3);
                if (dist <= rs_worst(&rs)) rs_add(&rs, dist, j);
            }
            int64_t *o = out_idx + ((size_t)b * N2 + i) * K;
            for (long j = 0; j < K; ++j) o[j] = j < rs.count ? ids[j] : 0;
            if (out_dist) {
                float *od = out_dist + ((size_t)b * N2 + i) * K;
                for (long j = 0; j < K; ++j) od[j] = j < rs.count ? ds[j] : FLT_MAX;
            }
        }
}

/* fp32 squared distances of given neighbour lists, with the reference arithmetic (for tie-group checks). */
void pu_oracle_knn_dists(const float *support, long B, long N1, const float *query, long N2, long K,
                         const int32_t *idx, float *out_dist)
{
#pragma omp parallel for
    for (long r = 0; r < B * N2; ++r) {
        const long b = r / N2;
        const float *q = query + (size_t)r * 3;
        for (long j = 0; j < K; ++j) {
            const long id = idx[(size_t)r * K + j];
            out_dist[(size_t)r * K + j] = l2_eval(q, support + ((size_t)b * N1 + id) * 3);
        }
    }
}
