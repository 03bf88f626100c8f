/*
 * petal_oracle_impl.h -- type-generic body of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C restatement of the algorithms in
 * petabi/petal-neighbors v0.18.0 (Rust).  It is included twice by petal_oracle.c, once with
 * REAL=float / SFX=f32 and once with REAL=double / SFX=f64 (the reference is generic over
 * `A: Float`).  Nothing under petal-neighbors_b200/ may include, link or call it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Every function cites the reference file:line it follows.  Compile with
 * `-O2 -ffp-contract=off` and no -ffast-math: Rust never contracts a*b+c into an FMA, so the
 * distance fold below has to stay a separate multiply and add.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

/* ------------------------------------------------------------------------------------------
 * L1 metric: Euclidean::distance, src/distance.rs:26-35.
 * Sequential left fold over the dimensions: diff = v1 - v2; sum += diff * diff; then sqrt.
 * ---------------------------------------------------------------------------------------- */
REAL FN(orc_distance)(const REAL *x1, const REAL *x2, size_t d)
{
    REAL sum = (REAL)0;
    for (size_t j = 0; j < d; ++j) {
        REAL diff = x1[j] - x2[j];
        sum += diff * diff;
    }
    return SQRT(sum);
}

/* Euclidean::rdistance, src/distance.rs:37-45 (no sqrt).  Dead on the tree paths; kept for the
 * pairwise / metric KATs. */
REAL FN(orc_rdistance)(const REAL *x1, const REAL *x2, size_t d)
{
    REAL sum = (REAL)0;
    for (size_t j = 0; j < d; ++j) {
        REAL diff = x1[j] - x2[j];
        sum += diff * diff;
    }
    return sum;
}

/* distance::pairwise, src/distance.rs:58-74: dense symmetric n x n matrix, zero diagonal. */
void FN(orc_pairwise)(const REAL *x, size_t n, size_t d, size_t row_stride, REAL *out)
{
    memset(out, 0, n * n * sizeof(REAL));
    if (n < 2) return;
    for (size_t i = 0; i < n; ++i)
        for (size_t j = i + 1; j < n; ++j) {
            REAL v = FN(orc_distance)(x + i * row_stride, x + j * row_stride, d);
            out[i * n + j] = v;
            out[j * n + i] = v;
        }
}

/* ------------------------------------------------------------------------------------------
 * Ball tree data layout: struct BallTree src/ball_tree.rs:15-24, struct Node :427-432.
 * Centroids are stored in one dense array instead of one heap Array1 per node.
 * ---------------------------------------------------------------------------------------- */
typedef struct FN(orc_balltree) {
    const REAL *points; /* borrowed, row-major, row stride `stride` (CowArray view) */
    size_t n, d, stride;
    size_t *idx;       /* BallTree::idx */
    size_t n_nodes;    /* 2^height - 1, src/ball_tree.rs:51-52 */
    size_t *range_lo, *range_hi;
    REAL *centroid;    /* n_nodes x d */
    REAL *radius;
    unsigned char *is_leaf;
} FN(orc_balltree);

/* Node::init, src/ball_tree.rs:445-461: centroid = (sum of rows in idx order) / len,
 * radius = max over idx of distance(centroid, row). */
void FN(orc_node_init)(const REAL *points, size_t d, size_t stride, const size_t *idx,
                       size_t n_idx, REAL *centroid_out, REAL *radius_out)
{
    for (size_t j = 0; j < d; ++j) centroid_out[j] = (REAL)0;
    for (size_t t = 0; t < n_idx; ++t) {
        const REAL *row = points + idx[t] * stride;
        for (size_t j = 0; j < d; ++j) centroid_out[j] += row[j];
    }
    REAL len = (REAL)n_idx; /* A::from_usize(idx.len()) */
    for (size_t j = 0; j < d; ++j) centroid_out[j] /= len;
    REAL mx = (REAL)0;
    for (size_t t = 0; t < n_idx; ++t) {
        REAL v = FN(orc_distance)(centroid_out, points + idx[t] * stride, d);
        if (v > mx) mx = v; /* A::max(dist, max) */
    }
    *radius_out = mx;
}

/* max_spread_column, src/ball_tree.rs:577-613: per column min/max over idx; the first column
 * with the strictly greatest spread wins (:605 `Some(Ordering::Greater)`).
 * Returns (size_t)-1 for the reference's "empty matrix" panic. */
size_t FN(orc_max_spread_column)(const REAL *points, size_t n_rows, size_t d, size_t stride,
                                 const size_t *idx, size_t n_idx)
{
    if (d == 0 || n_idx == 0) return (size_t)-1;
    for (size_t t = 0; t < n_idx; ++t)
        if (idx[t] >= n_rows) return (size_t)-2; /* "index out of bounds" panic :583-586 */
    size_t best_col = 0;
    REAL best = (REAL)0;
    for (size_t j = 0; j < d; ++j) {
        REAL mn = points[idx[0] * stride + j], mx = mn;
        for (size_t t = 1; t < n_idx; ++t) {
            REAL v = points[idx[t] * stride + j];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        REAL spread = mx - mn;
        if (j == 0) {
            best = spread;
        } else if (spread > best) {
            best = spread;
            best_col = j;
        }
    }
    return best_col;
}

/* halve_node_indices, src/ball_tree.rs:545-569: in-place Lomuto quick-select on one column,
 * pivot = last element, until the pivot lands on mid = len / 2.  Returns -1 for the
 * reference's empty-slice overflow panic (:549, test :800-806). */
int FN(orc_halve_node_indices)(size_t *idx, size_t n_idx, const REAL *col, size_t col_stride)
{
    if (n_idx == 0) return -1;
    size_t first = 0, last = n_idx - 1;
    size_t mid = n_idx / 2;
    for (;;) {
        size_t cur = first;
        for (size_t i = first; i < last; ++i) {
            if (col[idx[i] * col_stride] < col[idx[last] * col_stride]) {
                size_t t = idx[i]; idx[i] = idx[cur]; idx[cur] = t;
                cur += 1;
            }
        }
        { size_t t = idx[cur]; idx[cur] = idx[last]; idx[last] = t; }
        if (cur == mid) break;
        if (cur < mid) first = cur + 1; else last = cur - 1;
    }
    return 0;
}

/* build_subtree, src/ball_tree.rs:504-538: pre-order recursion, leaf iff 2*root+1 >= n_nodes,
 * split on the max-spread column at mid = (start + end) / 2. */
static void FN(orc_build_subtree)(FN(orc_balltree) *t, size_t root, size_t lo, size_t hi)
{
    FN(orc_node_init)(t->points, t->d, t->stride, t->idx + lo, hi - lo,
                      t->centroid + root * t->d, &t->radius[root]);
    t->range_lo[root] = lo;
    t->range_hi[root] = hi;
    size_t left = root * 2 + 1;
    if (left >= t->n_nodes) {
        t->is_leaf[root] = 1;
        return;
    }
    size_t col = FN(orc_max_spread_column)(t->points, t->n, t->d, t->stride, t->idx + lo, hi - lo);
    FN(orc_halve_node_indices)(t->idx + lo, hi - lo, t->points + col, t->stride);
    size_t mid = (lo + hi) / 2;
    FN(orc_build_subtree)(t, left, lo, mid);
    FN(orc_build_subtree)(t, left + 1, mid, hi);
}

/* BallTree::new / ::euclidean, src/ball_tree.rs:38-63, 367-373.
 * err: 0 ok, 1 ArrayError::Empty (:44-46), 2 ArrayError::NotContiguous (:47-49: row 0 must
 * have unit element stride; the row-to-row stride is free). */
FN(orc_balltree) *FN(orc_balltree_new)(const REAL *points, size_t n, size_t d, size_t row_stride,
                                       size_t col_stride, int *err)
{
    *err = 0;
    if (n == 0) { *err = 1; return NULL; }
    if (col_stride != 1 && d > 1) { *err = 2; return NULL; }
    FN(orc_balltree) *t = (FN(orc_balltree) *)calloc(1, sizeof(*t));
    t->points = points; t->n = n; t->d = d; t->stride = row_stride;
    unsigned height = 0;
    for (size_t v = n; v; v >>= 1) ++height; /* usize::BITS - leading_zeros */
    t->n_nodes = ((size_t)1 << height) - 1;
    t->idx = (size_t *)malloc(n * sizeof(size_t));
    for (size_t i = 0; i < n; ++i) t->idx[i] = i;
    t->range_lo = (size_t *)calloc(t->n_nodes, sizeof(size_t));
    t->range_hi = (size_t *)calloc(t->n_nodes, sizeof(size_t));
    t->centroid = (REAL *)calloc(t->n_nodes * (d ? d : 1), sizeof(REAL));
    t->radius = (REAL *)calloc(t->n_nodes, sizeof(REAL));
    t->is_leaf = (unsigned char *)calloc(t->n_nodes, 1);
    FN(orc_build_subtree)(t, 0, 0, n);
    return t;
}

void FN(orc_balltree_free)(FN(orc_balltree) *t)
{
    if (!t) return;
    free(t->idx); free(t->range_lo); free(t->range_hi);
    free(t->centroid); free(t->radius); free(t->is_leaf); free(t);
}

/* accessors: num_nodes :346-348, num_points :351-353, points_of :331-333, radius_of :336-338,
 * children_of :320-328 */
size_t FN(orc_balltree_num_nodes)(const FN(orc_balltree) *t) { return t->n_nodes; }
size_t FN(orc_balltree_num_points)(const FN(orc_balltree) *t) { return t->n; }
const size_t *FN(orc_balltree_idx)(const FN(orc_balltree) *t) { return t->idx; }
void FN(orc_balltree_node)(const FN(orc_balltree) *t, size_t node, size_t *lo, size_t *hi,
                           REAL *radius, int *is_leaf, REAL *centroid_out)
{
    *lo = t->range_lo[node]; *hi = t->range_hi[node];
    *radius = t->radius[node]; *is_leaf = t->is_leaf[node];
    if (centroid_out) memcpy(centroid_out, t->centroid + node * t->d, t->d * sizeof(REAL));
}

/* Node::distance_lower_bound, src/ball_tree.rs:473-481 */
static inline REAL FN(orc_lower_bound)(const FN(orc_balltree) *t, size_t node, const REAL *q,
                                       unsigned long long *n_dist)
{
    if (n_dist) ++*n_dist;
    REAL cd = FN(orc_distance)(q, t->centroid + node * t->d, t->d);
    REAL lb = cd - t->radius[node];
    return lb < (REAL)0 ? (REAL)0 : lb;
}

/* node_distance_lower_bound, src/ball_tree.rs:303-317 */
REAL FN(orc_balltree_node_distance_lower_bound)(const FN(orc_balltree) *t, size_t n1, size_t n2)
{
    REAL lb = FN(orc_distance)(t->centroid + n1 * t->d, t->centroid + n2 * t->d, t->d)
              - t->radius[n1] - t->radius[n2];
    return lb < (REAL)0 ? (REAL)0 : lb;
}

/* nearest_neighbor_in_subtree, src/ball_tree.rs:149-196.  Returns 1 = Some, 0 = None. */
int FN(orc_balltree_nearest_in_subtree)(const FN(orc_balltree) *t, const REAL *q, size_t root,
                                        REAL radius, size_t *out_i, REAL *out_d)
{
    REAL lower = FN(orc_lower_bound)(t, root, q, NULL);
    if (lower > radius) return 0;
    if (t->is_leaf[root]) {
        size_t min_i = 0;
        REAL min_dist = (REAL)INFINITY;
        for (size_t p = t->range_lo[root]; p < t->range_hi[root]; ++p) {
            size_t i = t->idx[p];
            REAL dist = FN(orc_distance)(q, t->points + i * t->stride, t->d);
            if (dist < min_dist) { min_i = i; min_dist = dist; }
        }
        if (min_dist <= radius) { *out_i = min_i; *out_d = min_dist; return 1; }
        return 0;
    }
    size_t c1 = root * 2 + 1, c2 = c1 + 1;
    REAL lb1 = FN(orc_lower_bound)(t, c1, q, NULL);
    REAL lb2 = FN(orc_lower_bound)(t, c2, q, NULL);
    if (!(lb1 < lb2)) { size_t s = c1; c1 = c2; c2 = s; }
    size_t i1; REAL d1;
    if (FN(orc_balltree_nearest_in_subtree)(t, q, c1, radius, &i1, &d1)) {
        size_t i2; REAL d2;
        /* .map_or(Some(neighbor), Some): the second child's answer overrides, also on a tie */
        if (FN(orc_balltree_nearest_in_subtree)(t, q, c2, d1, &i2, &d2)) { *out_i = i2; *out_d = d2; }
        else { *out_i = i1; *out_d = d1; }
        return 1;
    }
    return FN(orc_balltree_nearest_in_subtree)(t, q, c2, radius, out_i, out_d);
}

/* BallTree::query_nearest, src/ball_tree.rs:80-86 */
void FN(orc_balltree_query_nearest)(const FN(orc_balltree) *t, const REAL *q, size_t *out_i,
                                    REAL *out_d)
{
    int some = FN(orc_balltree_nearest_in_subtree)(t, q, 0, (REAL)INFINITY, out_i, out_d);
    if (!some) { *out_i = (size_t)-1; *out_d = (REAL)NAN; } /* .expect() would panic */
}

/* struct Neighbor + Ord, src/ball_tree.rs:378-423: ordered by distance only (OrderedFloat).
 * The max-heap below restates Rust std `alloc::collections::BinaryHeap` (push = sift_up,
 * pop = swap-with-last + sift_down_to_bottom + sift_up, into_sorted_vec = repeated
 * swap(0,end) + sift_down_range).  std is not part of /root/reference; its order among
 * EQUAL distances is therefore a best-effort restatement (unpinned by any test). */
typedef struct { size_t idx; REAL dist; } FN(orc_nb);
typedef struct { FN(orc_nb) *a; size_t len; } FN(orc_heap);

static void FN(heap_sift_up)(FN(orc_nb) *a, size_t start, size_t pos)
{
    FN(orc_nb) e = a[pos];
    while (pos > start) {
        size_t parent = (pos - 1) / 2;
        if (e.dist <= a[parent].dist) break;
        a[pos] = a[parent];
        pos = parent;
    }
    a[pos] = e;
}
static void FN(heap_sift_down_range)(FN(orc_nb) *a, size_t pos, size_t end)
{
    FN(orc_nb) e = a[pos];
    size_t child = 2 * pos + 1;
    while (end >= 2 && child <= end - 2) {
        if (a[child].dist <= a[child + 1].dist) child += 1;
        if (e.dist >= a[child].dist) { a[pos] = e; return; }
        a[pos] = a[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (child == end - 1 && e.dist < a[child].dist) { a[pos] = a[child]; pos = child; }
    a[pos] = e;
}
static void FN(heap_sift_down_to_bottom)(FN(orc_nb) *a, size_t pos, size_t end)
{
    size_t start = pos;
    FN(orc_nb) e = a[pos];
    size_t child = 2 * pos + 1;
    while (end >= 2 && child <= end - 2) {
        if (a[child].dist <= a[child + 1].dist) child += 1;
        a[pos] = a[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (child == end - 1) { a[pos] = a[child]; pos = child; }
    a[pos] = e;
    FN(heap_sift_up)(a, start, pos);
}
static void FN(heap_push)(FN(orc_heap) *h, FN(orc_nb) e)
{
    h->a[h->len] = e;
    FN(heap_sift_up)(h->a, 0, h->len);
    h->len += 1;
}
static void FN(heap_pop)(FN(orc_heap) *h)
{
    h->len -= 1;
    if (h->len > 0) {
        FN(orc_nb) last = h->a[h->len];
        h->a[h->len] = h->a[0];
        h->a[0] = last;
        FN(heap_sift_down_to_bottom)(h->a, 0, h->len);
    }
}
static void FN(heap_into_sorted)(FN(orc_heap) *h)
{
    size_t end = h->len;
    while (end > 1) {
        end -= 1;
        FN(orc_nb) t = h->a[0]; h->a[0] = h->a[end]; h->a[end] = t;
        FN(heap_sift_down_range)(h->a, 0, end);
    }
}

/* nearest_k_neighbors_in_subtree, src/ball_tree.rs:203-243 */
static void FN(orc_knn_subtree)(const FN(orc_balltree) *t, const REAL *q, size_t root,
                                REAL *radius, size_t k, FN(orc_heap) *h,
                                unsigned long long *n_dist)
{
    if (FN(orc_lower_bound)(t, root, q, n_dist) > *radius) return;
    if (t->is_leaf[root]) {
        for (size_t p = t->range_lo[root]; p < t->range_hi[root]; ++p) {
            size_t i = t->idx[p];
            if (n_dist) ++*n_dist;
            FN(orc_nb) nb = { i, FN(orc_distance)(q, t->points + i * t->stride, t->d) };
            if (h->len < k) {
                FN(heap_push)(h, nb);
            } else if (nb.dist < h->a[0].dist) {
                FN(heap_pop)(h);
                FN(heap_push)(h, nb);
            }
        }
    } else {
        size_t c1 = root * 2 + 1, c2 = c1 + 1;
        REAL lb1 = FN(orc_lower_bound)(t, c1, q, n_dist);
        REAL lb2 = FN(orc_lower_bound)(t, c2, q, n_dist);
        if (!(lb1 < lb2)) { size_t s = c1; c1 = c2; c2 = s; }
        FN(orc_knn_subtree)(t, q, c1, radius, k, h, n_dist);
        FN(orc_knn_subtree)(t, q, c2, radius, k, h, n_dist);
    }
    if (h->len == k) *radius = h->a[0].dist;
}

/* BallTree::query, src/ball_tree.rs:102-121.  Returns the result length (min(k, n); 0 for
 * k == 0), ascending by distance.  n_dist (optional) counts metric.distance evaluations. */
size_t FN(orc_balltree_query)(const FN(orc_balltree) *t, const REAL *q, size_t k, size_t *out_i,
                              REAL *out_d, unsigned long long *n_dist)
{
    if (k == 0) return 0;
    FN(orc_heap) h;
    size_t cap = k < t->n ? k : t->n; /* with_capacity(k); never holds more than n */
    h.a = (FN(orc_nb) *)malloc((cap + 1) * sizeof(FN(orc_nb)));
    h.len = 0;
    REAL radius = (REAL)INFINITY;
    FN(orc_knn_subtree)(t, q, 0, &radius, k, &h, n_dist);
    FN(heap_into_sorted)(&h);
    for (size_t i = 0; i < h.len; ++i) { out_i[i] = h.a[i].idx; out_d[i] = h.a[i].dist; }
    size_t len = h.len;
    free(h.a);
    return len;
}

/* neighbors_within_radius_in_subtree, src/ball_tree.rs:250-294 (BallTree::query_radius
 * :137-142): explicit stack, right child popped first, bulk include when ub <= r with no
 * per-point test, strict dist < r in leaves.  Caller frees *out with orc_free. */
size_t FN(orc_balltree_query_radius)(const FN(orc_balltree) *t, const REAL *q, REAL r,
                                     size_t **out)
{
    size_t cap = 16, len = 0;
    size_t *res = (size_t *)malloc(cap * sizeof(size_t));
    size_t scap = 64, sp = 0;
    size_t *stack = (size_t *)malloc(scap * sizeof(size_t));
    stack[sp++] = 0;
    while (sp) {
        size_t node = stack[--sp];
        REAL cd = FN(orc_distance)(q, t->centroid + node * t->d, t->d); /* distance_bounds :463-471 */
        REAL lb = cd - t->radius[node];
        if (lb < (REAL)0) lb = (REAL)0;
        REAL ub = cd + t->radius[node];
        if (lb > r) continue;
        size_t lo = t->range_lo[node], hi = t->range_hi[node];
        if (ub <= r) {
            if (len + (hi - lo) > cap) { while (len + (hi - lo) > cap) cap *= 2; res = (size_t *)realloc(res, cap * sizeof(size_t)); }
            for (size_t p = lo; p < hi; ++p) res[len++] = t->idx[p];
        } else if (t->is_leaf[node]) {
            for (size_t p = lo; p < hi; ++p) {
                size_t i = t->idx[p];
                REAL dist = FN(orc_distance)(q, t->points + i * t->stride, t->d);
                if (dist < r) {
                    if (len == cap) { cap *= 2; res = (size_t *)realloc(res, cap * sizeof(size_t)); }
                    res[len++] = i;
                }
            }
        } else {
            if (sp + 2 > scap) { scap *= 2; stack = (size_t *)realloc(stack, scap * sizeof(size_t)); }
            stack[sp++] = node * 2 + 1;
            stack[sp++] = node * 2 + 2;
        }
    }
    free(stack);
    *out = res;
    return len;
}

/* ------------------------------------------------------------------------------------------
 * Vantage-point tree, src/vantage_point_tree.rs.
 * ---------------------------------------------------------------------------------------- */
typedef struct { REAL dist; size_t id; size_t pos; } FN(orc_di); /* DistanceIndex :209-212 */

typedef struct FN(orc_vptree) {
    const REAL *points;
    size_t n, d, stride;
    size_t n_nodes;
    size_t *near_, *far_, *vp; /* Node :200-205 */
    REAL *mu;                  /* Node::radius */
    size_t root;
} FN(orc_vptree);

#define ORC_NULL ((size_t)-1) /* const NULL: usize = usize::MAX :207 */

static int FN(orc_di_cmp)(const void *a, const void *b)
{
    const FN(orc_di) *x = (const FN(orc_di) *)a, *y = (const FN(orc_di) *)b;
    if (x->dist < y->dist) return -1;
    if (x->dist > y->dist) return 1;
    /* sort_unstable_by_key (:178) may leave equal keys in any order; a stable order
     * (by position in the slice) is one legal outcome and makes the oracle deterministic. */
    return (x->pos > y->pos) - (x->pos < y->pos);
}

/* create_node, src/vantage_point_tree.rs:146-197 */
static size_t FN(orc_vp_create_node)(FN(orc_vptree) *t, FN(orc_di) *ix, size_t len)
{
    if (len == 0) return ORC_NULL;
    if (len == 1) {
        size_t id = t->n_nodes++;
        t->near_[id] = ORC_NULL; t->far_[id] = ORC_NULL;
        t->vp[id] = ix[0].id;
        t->mu[id] = REAL_MAX; /* A::max_value() :165 */
        return id;
    }
    size_t vp_pos = len - 1;
    size_t vantage = ix[vp_pos].id;
    for (size_t r = 0; r < vp_pos; ++r) {
        ix[r].dist = FN(orc_distance)(t->points + ix[r].id * t->stride,
                                      t->points + vantage * t->stride, t->d);
        ix[r].pos = r;
    }
    qsort(ix, vp_pos, sizeof(FN(orc_di)), FN(orc_di_cmp));
    size_t half = vp_pos / 2;
    REAL radius = ix[half].dist; /* far[0].distance :182 */
    size_t id = t->n_nodes++;
    t->near_[id] = ORC_NULL; t->far_[id] = ORC_NULL;
    t->vp[id] = vantage;
    t->mu[id] = radius;
    size_t nr = FN(orc_vp_create_node)(t, ix, half);
    size_t fr = FN(orc_vp_create_node)(t, ix + half, vp_pos - half);
    t->near_[id] = nr;
    t->far_[id] = fr;
    return id;
}

/* VantagePointTree::new / ::euclidean, src/vantage_point_tree.rs:31-72, create_root :132-144 */
FN(orc_vptree) *FN(orc_vptree_new)(const REAL *points, size_t n, size_t d, size_t row_stride,
                                   size_t col_stride, int *err)
{
    *err = 0;
    if (n == 0) { *err = 1; return NULL; }
    if (col_stride != 1 && d > 1) { *err = 2; return NULL; }
    FN(orc_vptree) *t = (FN(orc_vptree) *)calloc(1, sizeof(*t));
    t->points = points; t->n = n; t->d = d; t->stride = row_stride;
    t->near_ = (size_t *)malloc(n * sizeof(size_t));
    t->far_ = (size_t *)malloc(n * sizeof(size_t));
    t->vp = (size_t *)malloc(n * sizeof(size_t));
    t->mu = (REAL *)malloc(n * sizeof(REAL));
    FN(orc_di) *ix = (FN(orc_di) *)malloc(n * sizeof(FN(orc_di)));
    for (size_t i = 0; i < n; ++i) { ix[i].dist = REAL_MAX; ix[i].id = i; ix[i].pos = i; }
    t->n_nodes = 0;
    t->root = FN(orc_vp_create_node)(t, ix, n);
    free(ix);
    return t;
}

void FN(orc_vptree_free)(FN(orc_vptree) *t)
{
    if (!t) return;
    free(t->near_); free(t->far_); free(t->vp); free(t->mu); free(t);
}

/* search_node, src/vantage_point_tree.rs:100-130 */
static void FN(orc_vp_search)(const FN(orc_vptree) *t, size_t node, const REAL *q, REAL *best_d,
                              size_t *best_i, unsigned long long *n_dist)
{
    if (n_dist) ++*n_dist;
    REAL dist = FN(orc_distance)(t->points + t->vp[node] * t->stride, q, t->d);
    if (dist < *best_d) { *best_d = dist; *best_i = t->vp[node]; }
    size_t nr = t->near_[node], fr = t->far_[node];
    if (dist < t->mu[node]) {
        if (nr != ORC_NULL) FN(orc_vp_search)(t, nr, q, best_d, best_i, n_dist);
        if (fr != ORC_NULL && dist + *best_d > t->mu[node])
            FN(orc_vp_search)(t, fr, q, best_d, best_i, n_dist);
    } else {
        if (fr != ORC_NULL) FN(orc_vp_search)(t, fr, q, best_d, best_i, n_dist);
        if (nr != ORC_NULL && dist - *best_d < t->mu[node])
            FN(orc_vp_search)(t, nr, q, best_d, best_i, n_dist);
    }
}

/* VantagePointTree::query_nearest, src/vantage_point_tree.rs:88-98 */
void FN(orc_vptree_query_nearest)(const FN(orc_vptree) *t, const REAL *q, size_t *out_i,
                                  REAL *out_d, unsigned long long *n_dist)
{
    REAL best_d = REAL_MAX;
    size_t best_i = ORC_NULL;
    FN(orc_vp_search)(t, t->root, q, &best_d, &best_i, n_dist);
    *out_i = best_i;
    *out_d = best_d;
}

/* ------------------------------------------------------------------------------------------
 * Brute-force oracle: the tie-break authority.  Sanctioned by the reference's own
 * naive_k_nearest_neighbors, src/ball_tree.rs:873-894 (all distances, sort, take k), made
 * total by ordering on (distance, index).  Output rows are padded with (SIZE_MAX, +inf)
 * when k > n.
 * ---------------------------------------------------------------------------------------- */
void FN(orc_brute_knn)(const REAL *points, size_t n, size_t d, size_t stride, const REAL *q,
                       size_t nq, size_t q_stride, size_t k, size_t *out_i, REAL *out_d,
                       int n_threads)
{
    if (k == 0) return;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#endif
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long long qi = 0; qi < (long long)nq; ++qi) {
        size_t *bi = out_i + (size_t)qi * k;
        REAL *bd = out_d + (size_t)qi * k;
        size_t len = 0;
        for (size_t i = 0; i < n; ++i) {
            REAL dist = FN(orc_distance)(q + (size_t)qi * q_stride, points + i * stride, d);
            if (len == k && !(dist < bd[k - 1])) continue; /* i ascending: ties keep the lower index */
            size_t pos = len < k ? len : k - 1;
            while (pos > 0 && dist < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
            bd[pos] = dist; bi[pos] = i;
            if (len < k) ++len;
        }
        for (size_t s = len; s < k; ++s) { bi[s] = (size_t)-1; bd[s] = (REAL)INFINITY; }
    }
}

/* Brute-force radius oracle: indices with distance < r (strict, src/ball_tree.rs:277),
 * ascending by index.  Two calls: counts first (out == NULL), then fill with offsets. */
void FN(orc_brute_radius)(const REAL *points, size_t n, size_t d, size_t stride, const REAL *q,
                          size_t nq, size_t q_stride, REAL r, size_t *counts,
                          const size_t *offsets, size_t *out, int n_threads)
{
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#endif
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long long qi = 0; qi < (long long)nq; ++qi) {
        size_t c = 0;
        for (size_t i = 0; i < n; ++i) {
            REAL dist = FN(orc_distance)(q + (size_t)qi * q_stride, points + i * stride, d);
            if (dist < r) {
                if (out) out[offsets[qi] + c] = i;
                ++c;
            }
        }
        if (counts) counts[qi] = c;
    }
}

/* ------------------------------------------------------------------------------------------
 * Batch drivers for the CPU baseline (bench.py cpu_baseline / --impl reference): the bench
 * loop of benches/ball_tree.rs:53-59 (one query per call) under an optional OpenMP outer loop
 * (the rayon-equivalent; legal because queries take &self and Euclidean is Sync,
 * src/distance.rs:19).  n_threads == 1 is the reference "as shipped".
 * ---------------------------------------------------------------------------------------- */
unsigned long long FN(orc_balltree_query_batch)(const FN(orc_balltree) *t, const REAL *q,
                                                size_t nq, size_t q_stride, size_t k,
                                                size_t *out_i, REAL *out_d, int n_threads)
{
    unsigned long long total = 0;
    size_t kk = k < t->n ? k : t->n;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#endif
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(+ : total)
    for (long long qi = 0; qi < (long long)nq; ++qi) {
        unsigned long long nd = 0;
        size_t len = FN(orc_balltree_query)(t, q + (size_t)qi * q_stride, k, out_i + (size_t)qi * k,
                                            out_d + (size_t)qi * k, &nd);
        for (size_t s = len; s < k; ++s) { out_i[(size_t)qi * k + s] = (size_t)-1; out_d[(size_t)qi * k + s] = (REAL)INFINITY; }
        (void)kk;
        total += nd;
    }
    return total;
}

unsigned long long FN(orc_vptree_query_nearest_batch)(const FN(orc_vptree) *t, const REAL *q,
                                                      size_t nq, size_t q_stride, size_t *out_i,
                                                      REAL *out_d, int n_threads)
{
    unsigned long long total = 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#endif
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(+ : total)
    for (long long qi = 0; qi < (long long)nq; ++qi) {
        unsigned long long nd = 0;
        FN(orc_vptree_query_nearest)(t, q + (size_t)qi * q_stride, &out_i[qi], &out_d[qi], &nd);
        total += nd;
    }
    return total;
}

/* Radius batch: counts[nq] always written; when out != NULL the hits of query qi are written
 * at out + offsets[qi] in the reference's DFS order. */
void FN(orc_balltree_query_radius_batch)(const FN(orc_balltree) *t, const REAL *q, size_t nq,
                                         size_t q_stride, REAL r, size_t *counts,
                                         const size_t *offsets, size_t *out, int n_threads)
{
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#endif
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long long qi = 0; qi < (long long)nq; ++qi) {
        size_t *res = NULL;
        size_t len = FN(orc_balltree_query_radius)(t, q + (size_t)qi * q_stride, r, &res);
        counts[qi] = len;
        if (out) memcpy(out + offsets[qi], res, len * sizeof(size_t));
        free(res);
    }
}

#undef FN
#undef CAT
#undef CAT_
