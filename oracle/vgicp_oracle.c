/*
 * vgicp_oracle.c -- CPU restatement of the fast_gicp VGICP-CUDA hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fast_gicp_b200/ may include, link or call this file; it is
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker.
 *
 * PARITY STATUS: the reference (koide3/fast_gicp) cannot be compiled in this environment (no Eigen, PCL,
 * Boost, FLANN; thirdparty/ submodules are empty), so this restatement is pinned only by what the reference's
 * own tests/fixtures hold for this path:
 *   - data/relative.txt end-to-end pose within 0.05 m / 1 deg under the four call orders of
 *     src/test/gicp_test.cpp:147-201 (tests/test_oracle_golden.py),
 *   - README.md:116 point counts for the downsampled benchmark inputs (tests/golden/make_fixtures.py).
 * Everything finer (per-stage values) is "parity unpinned": no golden vectors exist in the reference.
 * Eigen's SelfAdjointEigenSolver<Matrix3f>::computeDirect, Matrix3f::inverse(), LDLT and
 * Quaterniond::toRotationMatrix (Eigen @ 1fd5ce10, not vendored) are restated from the published algorithm.
 *
 * Float semantics: every device-side quantity of the reference is float32; this file is compiled with
 * -ffp-contract=off and spells out each fused multiply-add it wants with fmaf(), so that the CUDA path (which
 * spells out the same operations with __fmaf_rn/__fmul_rn/__fadd_rn) can be compared bit-for-bit where the
 * arithmetic is order-defined (voxel coordinates, hashes, kNN distances, raw covariances, transformed points).
 *
 * Where the reference's result depends on GPU thread arrival order (hash-slot ownership, voxel ids, float
 * atomicAdd order, thrust reduction tree) the oracle fixes one admissible order and says so at the function.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------------------
 * (1) hash + voxel coordinate  --  include/fast_gicp/cuda/vector3_hash.cuh:8-38
 * ---------------------------------------------------------------------------------------------------------- */
static inline void hash_combine(uint64_t* h, uint64_t k) { /* vector3_hash.cuh:8-20 */
  const uint64_t m = UINT64_C(0xc6a4a7935bd1e995);
  const int r = 47;
  k *= m;
  k ^= k >> r;
  k *= m;
  *h ^= k;
  *h *= m;
  *h += 0xe6546b64;
}

/* vector3_hash.cuh:27-33 -- the int argument converts to uint64_t by sign extension */
ORC_API uint64_t orc_vector3i_hash(int x, int y, int z) {
  uint64_t seed = 0;
  hash_combine(&seed, (uint64_t)(int64_t)x);
  hash_combine(&seed, (uint64_t)(int64_t)y);
  hash_combine(&seed, (uint64_t)(int64_t)z);
  return seed;
}

/* vector3_hash.cuh:35-38 -- (x.array() / resolution - 0.5).floor().cast<int>() in float */
static inline int voxel_coord1(float x, float res) { return (int)floorf(x / res - 0.5f); }
ORC_API void orc_voxel_coord(const float* p, float res, int* c) {
  c[0] = voxel_coord1(p[0], res);
  c[1] = voxel_coord1(p[1], res);
  c[2] = voxel_coord1(p[2], res);
}
ORC_API void orc_voxel_coords(const float* pts, int n, float res, int* coords) {
  for (int i = 0; i < n; i++) orc_voxel_coord(pts + 3 * i, res, coords + 3 * i);
}
ORC_API void orc_hashes(const int* coords, int n, uint64_t* out) {
  for (int i = 0; i < n; i++) out[i] = orc_vector3i_hash(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]);
}

/* ------------------------------------------------------------------------------------------------------------
 * (2) exact k-NN (self included)
 *     fast_vgicp_cuda_impl.hpp:152-167 (pcl::search::KdTree::nearestKSearch, default) and
 *     brute_force_knn.cu:16-60 (GPU_BRUTEFORCE) return the same SET; order differs (kd-tree: ascending distance,
 *     brute force: heap order).  The oracle returns ascending (d2, index) -- the kd-tree order with ties broken by
 *     index.  d2 = (dx*dx + dy*dy) + dz*dz in float, no contraction (FLANN L2_Simple / Eigen squaredNorm order).
 * ---------------------------------------------------------------------------------------------------------- */
static inline float sqdist3(const float* a, const float* b) {
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return (dx * dx + dy * dy) + dz * dz;
}

static inline int knn_less(float d, int i, float d2, int i2) { return d < d2 || (d == d2 && i < i2); }

/* insert (d,i) into ascending list of length *cnt<=k */
static inline void knn_insert(float* bd, int* bi, int* cnt, int k, float d, int i) {
  int n = *cnt;
  if (n == k) {
    if (!knn_less(d, i, bd[k - 1], bi[k - 1])) return;
    n = k - 1;
  }
  int p = n;
  while (p > 0 && knn_less(d, i, bd[p - 1], bi[p - 1])) {
    bd[p] = bd[p - 1];
    bi[p] = bi[p - 1];
    p--;
  }
  bd[p] = d;
  bi[p] = i;
  *cnt = n + 1;
}

ORC_API int orc_knn_bruteforce(const float* pts, int n, int k, int* idx_out, float* d2_out) {
  if (k <= 0 || k > n || k > 256) return -1;
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; i++) {
    float bd[256];
    int bi[256];
    int cnt = 0;
    for (int j = 0; j < n; j++) knn_insert(bd, bi, &cnt, k, sqdist3(pts + 3 * j, pts + 3 * i), j);
    for (int j = 0; j < k; j++) {
      idx_out[(size_t)i * k + j] = bi[j];
      if (d2_out) d2_out[(size_t)i * k + j] = bd[j];
    }
  }
  return 0;
}

/* --- kd-tree (exact; used for larger N and for the CPU baseline, standing in for pcl::search::KdTree/FLANN) --- */
typedef struct {
  int lo, hi;      /* point range in perm (leaf) */
  int left, right; /* children (-1 = leaf) */
  int dim;
  float split;
} KdNode;
typedef struct {
  const float* pts;
  int n;
  int* perm;
  KdNode* nodes;
  int n_nodes, cap_nodes;
} KdTree;

static int kd_build_rec(KdTree* t, int lo, int hi) {
  int id = t->n_nodes++;
  KdNode* nd = &t->nodes[id];
  nd->lo = lo;
  nd->hi = hi;
  nd->left = nd->right = -1;
  if (hi - lo <= 12) return id;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = lo; i < hi; i++)
    for (int d = 0; d < 3; d++) {
      float v = t->pts[3 * t->perm[i] + d];
      if (v < mn[d]) mn[d] = v;
      if (v > mx[d]) mx[d] = v;
    }
  int dim = 0;
  if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
  if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
  if (!(mx[dim] > mn[dim])) return id; /* all identical */
  /* median by nth_element (quickselect) */
  int mid = (lo + hi) / 2, l = lo, r = hi - 1;
  while (l < r) {
    float pv = t->pts[3 * t->perm[(l + r) / 2] + dim];
    int i = l, j = r;
    while (i <= j) {
      while (t->pts[3 * t->perm[i] + dim] < pv) i++;
      while (t->pts[3 * t->perm[j] + dim] > pv) j--;
      if (i <= j) {
        int tmp = t->perm[i];
        t->perm[i] = t->perm[j];
        t->perm[j] = tmp;
        i++;
        j--;
      }
    }
    if (mid <= j) r = j;
    else if (mid >= i) l = i;
    else break;
  }
  float split = t->pts[3 * t->perm[mid] + dim];
  t->nodes[id].dim = dim;
  t->nodes[id].split = split;
  int left = kd_build_rec(t, lo, mid);
  int right = kd_build_rec(t, mid, hi);
  t->nodes[id].left = left;
  t->nodes[id].right = right;
  return id;
}

static KdTree* kd_build(const float* pts, int n) {
  KdTree* t = (KdTree*)calloc(1, sizeof(KdTree));
  t->pts = pts;
  t->n = n;
  t->perm = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) t->perm[i] = i;
  t->cap_nodes = 2 * n + 2;
  t->nodes = (KdNode*)malloc(sizeof(KdNode) * (size_t)t->cap_nodes);
  t->n_nodes = 0;
  if (n > 0) kd_build_rec(t, 0, n);
  return t;
}
static void kd_free(KdTree* t) {
  if (!t) return;
  free(t->perm);
  free(t->nodes);
  free(t);
}

static void kd_search(const KdTree* t, int id, const float* q, int k, float* bd, int* bi, int* cnt) {
  const KdNode* nd = &t->nodes[id];
  if (nd->left < 0) {
    for (int i = nd->lo; i < nd->hi; i++) {
      int j = t->perm[i];
      knn_insert(bd, bi, cnt, k, sqdist3(t->pts + 3 * j, q), j);
    }
    return;
  }
  float diff = q[nd->dim] - nd->split;
  int first = diff < 0 ? nd->left : nd->right;
  int second = diff < 0 ? nd->right : nd->left;
  kd_search(t, first, q, k, bd, bi, cnt);
  /* <= so that equal-distance candidates with a smaller index are still seen (exact under ties) */
  if (*cnt < k || diff * diff <= bd[k - 1]) kd_search(t, second, q, k, bd, bi, cnt);
}

ORC_API int orc_knn_kdtree(const float* pts, int n, int k, int* idx_out) {
  if (k <= 0 || k > n || k > 256) return -1;
  KdTree* t = kd_build(pts, n);
#pragma omp parallel for schedule(guided, 8)
  for (int i = 0; i < n; i++) {
    float bd[256];
    int bi[256];
    int cnt = 0;
    kd_search(t, 0, pts + 3 * i, k, bd, bi, &cnt);
    for (int j = 0; j < k; j++) idx_out[(size_t)i * k + j] = bi[j];
  }
  kd_free(t);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * (3) covariance estimation  --  covariance_estimation.cu:26-34 (float, single pass, uncentred)
 *     mean += pt ; cov += pt*pt^T ; mean /= k ; cov = cov/k - mean*mean^T     (cov stored column-major 3x3)
 *     The multiply-adds are spelled fmaf (nvcc contracts them by default); the CUDA path spells the same.
 * ---------------------------------------------------------------------------------------------------------- */
ORC_API void orc_covariances(const float* pts, int n, int k, const int* nbr, float* cov9) {
#pragma omp parallel for schedule(static)
  for (int idx = 0; idx < n; idx++) {
    float mean[3] = {0.f, 0.f, 0.f};
    float c[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < k; i++) {
      const float* p = pts + 3 * (size_t)nbr[(size_t)idx * k + i];
      for (int d = 0; d < 3; d++) mean[d] = mean[d] + p[d];
      for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) c[col * 3 + row] = fmaf(p[row], p[col], c[col * 3 + row]);
    }
    float kf = (float)k;
    for (int d = 0; d < 3; d++) mean[d] = mean[d] / kf;
    for (int col = 0; col < 3; col++)
      for (int row = 0; row < 3; row++) cov9[(size_t)idx * 9 + col * 3 + row] = fmaf(-mean[row], mean[col], c[col * 3 + row] / kf);
  }
}

/* ------------------------------------------------------------------------------------------------------------
 * (3b) kernel-weighted covariance estimation (NearestNeighborMethod::GPU_RBF_KERNEL)  --  covariance_estimation_rbf.cu
 *     covariance_estimation_kernel :59-90 : for each block of BLOCK_SIZE = 512 consecutive points (the cloud is padded to a
 *       multiple of 512 with points at the origin, :126-129) and each query x: a NormalDistribution partial
 *       {sum_w, sum w*p, sum w*p*p^T} over the block's points with |x - p|^2 <= max_dist^2, w = exp(-kernel_width * |x - p|^2),
 *       accumulated in index order (NormalDistribution::accumulate :40-44);
 *     finalization_kernel :92-114 : the partials of a query are added in block order (operator+= :33-38), then
 *       NormalDistribution::finalize :46-52 : mean = sum_p / sum_w ; cov = (cov - mean * sum_p^T) / sum_w   (not symmetrised).
 *     A padding point inside max_dist of the query adds its weight to sum_w (and nothing else, it sits at the origin).
 *     float throughout, every a*b+c two roundings (the CUDA path's stage-1 unit is compiled --fmad=false and spells the same
 *     sequence).  The weight: the reference calls CUDA's expf (a 2-ulp function that no CPU libm reproduces bit for bit); both
 *     sides here evaluate the same exp built from IEEE single operations (exp_det, Cephes' expf: about 1 ulp) -- within the
 *     error class of the reference's own function.  cov9: column-major 3x3 per point like every other covariance of the oracle.
 * ---------------------------------------------------------------------------------------------------------- */
/* exp(x), x <= 0, from IEEE single operations only (Cephes' expf), the same sequence the CUDA path spells */
static float exp_det(float x) {
  if (x < -87.0f) return 0.0f;
  const float n = rintf(x * 1.44269504088896341f);
  float r = fmaf(n, -0.693359375f, x);
  r = fmaf(n, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = fmaf(p, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  p = fmaf(p, r * r, r) + 1.0f;
  union { int i; float f; } sc;
  sc.i = ((int)n + 127) << 23;
  return p * sc.f;
}

#define ORC_RBF_BLOCK 512
ORC_API void orc_covariances_rbf(const float* pts, int n, float kernel_width, float max_dist, float* cov9) {
  const float max_dist_sq = max_dist * max_dist;
  const int num_blocks = (n + ORC_RBF_BLOCK - 1) / ORC_RBF_BLOCK;
#pragma omp parallel for schedule(dynamic, 64)
  for (int q = 0; q < n; q++) {
    const float* x = pts + 3 * (size_t)q;
    float sw = 0.f, m[3] = {0.f, 0.f, 0.f}, c[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < num_blocks; b++) {
      float psw = 0.f, pm[3] = {0.f, 0.f, 0.f}, pc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < ORC_RBF_BLOCK; j++) {
        const int g = b * ORC_RBF_BLOCK + j;
        static const float origin[3] = {0.f, 0.f, 0.f};
        const float* p = g < n ? pts + 3 * (size_t)g : origin;
        const float dx = x[0] - p[0], dy = x[1] - p[1], dz = x[2] - p[2];
        const float sq = (dx * dx + dy * dy) + dz * dz;
        if (sq > max_dist_sq) continue;
        const float w = exp_det(-kernel_width * sq);
        psw += w;
        const float wp[3] = {w * p[0], w * p[1], w * p[2]};
        for (int d = 0; d < 3; d++) pm[d] += wp[d];
        for (int col = 0; col < 3; col++)
          for (int row = 0; row < 3; row++) pc[col * 3 + row] += wp[row] * p[col];
      }
      sw += psw;
      for (int d = 0; d < 3; d++) m[d] += pm[d];
      for (int d = 0; d < 9; d++) c[d] += pc[d];
    }
    const float mean[3] = {m[0] / sw, m[1] / sw, m[2] / sw};
    for (int col = 0; col < 3; col++)
      for (int row = 0; row < 3; row++) cov9[(size_t)q * 9 + col * 3 + row] = (c[col * 3 + row] - mean[row] * m[col]) / sw;
  }
}

/* ------------------------------------------------------------------------------------------------------------
 * (4) covariance regularisation  --  covariance_regularization.cu:15-25,34-52,74-101,105-124
 * ---------------------------------------------------------------------------------------------------------- */
/* column-major 3x3 helpers: M(r,c) = m[c*3+r] */
#define M3(m, r, c) ((m)[(c)*3 + (r)])

/* Eigen Matrix3f::inverse() (Eigen/src/LU/InverseImpl.h, compute_inverse<Matrix3,3>): cofactors / determinant,
 * determinant by expansion along column 0. */
static void inv3f(const float* m, float* out) {
  float c00 = M3(m, 1, 1) * M3(m, 2, 2) - M3(m, 1, 2) * M3(m, 2, 1);
  float c10 = M3(m, 2, 1) * M3(m, 0, 2) - M3(m, 2, 2) * M3(m, 0, 1); /* cofactor<1,0> */
  float c20 = M3(m, 0, 1) * M3(m, 1, 2) - M3(m, 0, 2) * M3(m, 1, 1); /* cofactor<2,0> */
  float det = (c00 * M3(m, 0, 0) + c10 * M3(m, 1, 0)) + c20 * M3(m, 2, 0);
  float invdet = 1.0f / det;
  /* result(j,i) = cofactor(i,j) * invdet */
  M3(out, 0, 0) = c00 * invdet;
  M3(out, 0, 1) = c10 * invdet;
  M3(out, 0, 2) = c20 * invdet;
  M3(out, 1, 0) = (M3(m, 1, 2) * M3(m, 2, 0) - M3(m, 1, 0) * M3(m, 2, 2)) * invdet; /* cofactor<0,1> */
  M3(out, 1, 1) = (M3(m, 2, 2) * M3(m, 0, 0) - M3(m, 2, 0) * M3(m, 0, 2)) * invdet; /* cofactor<1,1> */
  M3(out, 1, 2) = (M3(m, 0, 2) * M3(m, 1, 0) - M3(m, 0, 0) * M3(m, 1, 2)) * invdet; /* cofactor<2,1> */
  M3(out, 2, 0) = (M3(m, 1, 0) * M3(m, 2, 1) - M3(m, 1, 1) * M3(m, 2, 0)) * invdet; /* cofactor<0,2> */
  M3(out, 2, 1) = (M3(m, 2, 0) * M3(m, 0, 1) - M3(m, 2, 1) * M3(m, 0, 0)) * invdet; /* cofactor<1,2> */
  M3(out, 2, 2) = (M3(m, 0, 0) * M3(m, 1, 1) - M3(m, 0, 1) * M3(m, 1, 0)) * invdet; /* cofactor<2,2> */
}

static void mul3f(const float* a, const float* b, float* out) { /* out = a*b, column-major */
  float t[9];
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) t[c * 3 + r] = (M3(a, r, 0) * M3(b, 0, c) + M3(a, r, 1) * M3(b, 1, c)) + M3(a, r, 2) * M3(b, 2, c);
  memcpy(out, t, sizeof(t));
}

static void cross3f(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

/* Eigen/src/Eigenvalues/SelfAdjointEigenSolver.h direct_selfadjoint_eigenvalues<.,3,false>::computeRoots */
static void eig3_roots(const float* m, float* roots) {
  const float s_inv3 = 1.0f / 3.0f;
  const float s_sqrt3 = sqrtf(3.0f);
  float c0 = M3(m, 0, 0) * M3(m, 1, 1) * M3(m, 2, 2) + 2.0f * M3(m, 1, 0) * M3(m, 2, 0) * M3(m, 2, 1) - M3(m, 0, 0) * M3(m, 2, 1) * M3(m, 2, 1) -
             M3(m, 1, 1) * M3(m, 2, 0) * M3(m, 2, 0) - M3(m, 2, 2) * M3(m, 1, 0) * M3(m, 1, 0);
  float c1 = M3(m, 0, 0) * M3(m, 1, 1) - M3(m, 1, 0) * M3(m, 1, 0) + M3(m, 0, 0) * M3(m, 2, 2) - M3(m, 2, 0) * M3(m, 2, 0) + M3(m, 1, 1) * M3(m, 2, 2) -
             M3(m, 2, 1) * M3(m, 2, 1);
  float c2 = M3(m, 0, 0) + M3(m, 1, 1) + M3(m, 2, 2);
  float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c2 * c2_over_3 - c1) * s_inv3;
  if (a_over_3 < 0.0f) a_over_3 = 0.0f;
  float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = a_over_3 * a_over_3 * a_over_3 - half_b * half_b;
  if (q < 0.0f) q = 0.0f;
  float rho = sqrtf(a_over_3);
  /* atan2/cos/sin evaluated in double and rounded once to float: the correctly rounded float results, independent of
   * the libm underneath (Eigen calls the float functions of whatever platform it runs on: CUDA's are 2-ulp). The CUDA
   * path does the same, which makes the regularised covariances comparable bit-for-bit. */
  float theta = (float)atan2((double)sqrtf(q), (double)half_b) * s_inv3;
  float cos_theta = (float)cos((double)theta);
  float sin_theta = (float)sin((double)theta);
  roots[0] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[1] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 + 2.0f * rho * cos_theta;
}

/* extract_kernel: eigenvector = normalised larger cross product of the column with the largest |diagonal| with the
 * two other columns; that column is returned as "representative". */
static void eig3_extract_kernel(const float* mat, float* res, float* representative) {
  int i0 = 0;
  float best = fabsf(M3(mat, 0, 0));
  if (fabsf(M3(mat, 1, 1)) > best) { best = fabsf(M3(mat, 1, 1)); i0 = 1; }
  if (fabsf(M3(mat, 2, 2)) > best) { best = fabsf(M3(mat, 2, 2)); i0 = 2; }
  float rep[3] = {mat[i0 * 3 + 0], mat[i0 * 3 + 1], mat[i0 * 3 + 2]};
  float c0[3], c1[3];
  cross3f(rep, mat + ((i0 + 1) % 3) * 3, c0);
  cross3f(rep, mat + ((i0 + 2) % 3) * 3, c1);
  float n0 = (c0[0] * c0[0] + c0[1] * c0[1]) + c0[2] * c0[2];
  float n1 = (c1[0] * c1[0] + c1[1] * c1[1]) + c1[2] * c1[2];
  if (representative) { representative[0] = rep[0]; representative[1] = rep[1]; representative[2] = rep[2]; }
  if (n0 > n1) {
    float s = sqrtf(n0);
    res[0] = c0[0] / s; res[1] = c0[1] / s; res[2] = c0[2] / s;
  } else {
    float s = sqrtf(n1);
    res[0] = c1[0] / s; res[1] = c1[1] / s; res[2] = c1[2] / s;
  }
}

/* SelfAdjointEigenSolver<Matrix3f>::computeDirect -> eigenvalues ascending, eigenvectors in columns */
ORC_API void orc_eig3_direct(const float* cov, float* evals, float* evecs) {
  float s[9];
  /* selfadjointView<Lower>: mirror the lower triangle */
  M3(s, 0, 0) = M3(cov, 0, 0); M3(s, 1, 1) = M3(cov, 1, 1); M3(s, 2, 2) = M3(cov, 2, 2);
  M3(s, 1, 0) = M3(s, 0, 1) = M3(cov, 1, 0);
  M3(s, 2, 0) = M3(s, 0, 2) = M3(cov, 2, 0);
  M3(s, 2, 1) = M3(s, 1, 2) = M3(cov, 2, 1);
  float shift = ((M3(s, 0, 0) + M3(s, 1, 1)) + M3(s, 2, 2)) / 3.0f;
  M3(s, 0, 0) -= shift; M3(s, 1, 1) -= shift; M3(s, 2, 2) -= shift;
  float scale = 0.0f;
  for (int i = 0; i < 9; i++) if (fabsf(s[i]) > scale) scale = fabsf(s[i]);
  if (scale > 0.0f) for (int i = 0; i < 9; i++) s[i] /= scale;
  eig3_roots(s, evals);
  const float eps = 1.1920929e-07f;
  if ((evals[2] - evals[0]) <= eps) {
    for (int i = 0; i < 9; i++) evecs[i] = 0.0f;
    evecs[0] = evecs[4] = evecs[8] = 1.0f;
  } else {
    float tmp[9];
    memcpy(tmp, s, sizeof(tmp));
    float d0 = evals[2] - evals[1];
    float d1 = evals[1] - evals[0];
    int k = 0, l = 2;
    if (d0 > d1) { k = 2; l = 0; d0 = d1; }
    M3(tmp, 0, 0) -= evals[k]; M3(tmp, 1, 1) -= evals[k]; M3(tmp, 2, 2) -= evals[k];
    eig3_extract_kernel(tmp, evecs + 3 * k, evecs + 3 * l);
    if (d0 <= 2.0f * eps * d1) {
      float* vl = evecs + 3 * l; const float* vk = evecs + 3 * k;
      float dot = (vk[0] * vl[0] + vk[1] * vl[1]) + vk[2] * vl[2];
      float t0 = vl[0] - dot * vl[0], t1 = vl[1] - dot * vl[1], t2 = vl[2] - dot * vl[2];
      float nn = sqrtf((t0 * t0 + t1 * t1) + t2 * t2);
      if (nn > 0.0f) { vl[0] = t0 / nn; vl[1] = t1 / nn; vl[2] = t2 / nn; }
    } else {
      memcpy(tmp, s, sizeof(tmp));
      M3(tmp, 0, 0) -= evals[l]; M3(tmp, 1, 1) -= evals[l]; M3(tmp, 2, 2) -= evals[l];
      eig3_extract_kernel(tmp, evecs + 3 * l, NULL);
    }
    float c[3];
    cross3f(evecs + 6, evecs + 0, c);
    float nn = sqrtf((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
    if (nn > 0.0f) { evecs[3] = c[0] / nn; evecs[4] = c[1] / nn; evecs[5] = c[2] / nn; }
    else { evecs[3] = c[0]; evecs[4] = c[1]; evecs[5] = c[2]; }
  }
  for (int i = 0; i < 3; i++) evals[i] = evals[i] * scale + shift;
}

/* RegularizationMethod numeric order: gicp_settings.hpp:6  { NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS } */
enum { ORC_REG_NONE = 0, ORC_REG_MIN_EIG = 1, ORC_REG_NORMALIZED_MIN_EIG = 2, ORC_REG_PLANE = 3, ORC_REG_FROBENIUS = 4 };

static void reg_rebuild(const float* evecs, float l0, float l1, float l2, float* cov) { /* V * diag * V^-1 */
  float vinv[9], vd[9];
  inv3f(evecs, vinv);
  for (int r = 0; r < 3; r++) { M3(vd, r, 0) = M3(evecs, r, 0) * l0; M3(vd, r, 1) = M3(evecs, r, 1) * l1; M3(vd, r, 2) = M3(evecs, r, 2) * l2; }
  mul3f(vd, vinv, cov);
}

/* returns 0, or 1 when the method is not implemented on the reference's GPU path (covariances left untouched and
 * "unimplemented covariance regularization method" printed, covariance_regularization.cu:121-123) */
ORC_API int orc_regularize(float* cov9, int n, int method) {
  if (method == ORC_REG_NONE) return 1;           /* reference prints the message too; values untouched */
  if (method == ORC_REG_NORMALIZED_MIN_EIG) return 1;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    float* cov = cov9 + (size_t)i * 9;
    if (method == ORC_REG_PLANE) { /* :105-116 with svd_kernel :15-25, svd_reconstruction_kernel :34-52 */
      float ev[3], V[9];
      orc_eig3_direct(cov, ev, V);
      reg_rebuild(V, 1e-3f, 1.0f, 1.0f, cov);
    } else if (method == ORC_REG_MIN_EIG) { /* :84-101 */
      float ev[3], V[9];
      orc_eig3_direct(cov, ev, V);
      reg_rebuild(V, fmaxf(1e-3f, ev[0]), fmaxf(1e-3f, ev[1]), fmaxf(1e-3f, ev[2]), cov);
    } else if (method == ORC_REG_FROBENIUS) { /* :74-82 */
      float C[9], Ci[9];
      memcpy(C, cov, sizeof(C));
      C[0] += 1e-3f; C[4] += 1e-3f; C[8] += 1e-3f;
      inv3f(C, Ci);
      float nn = 0.0f;
      for (int j = 0; j < 9; j++) nn += Ci[j] * Ci[j];
      nn = sqrtf(nn);
      for (int j = 0; j < 9; j++) Ci[j] = Ci[j] / nn;
      inv3f(Ci, cov);
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * (5) Gaussian voxel map  --  gaussian_voxelmap.cu:21-58 (insert), :61-73, :76-120 (accumulate), :158-176 (finalize),
 *     :258-289 (table growth 8192*2^j until <1% of points fail to insert within 10 probes)
 *
 *     Arrival order: in the reference which voxel owns a contended slot, the voxel ids (atomicAdd) and the float
 *     atomicAdd order are thread-timing dependent.  The oracle (and the CUDA path) fix them as:
 *       - distinct voxels are inserted in lexicographic (x,y,z) order of their coordinate (an admissible arrival
 *         order; makes the table a pure function of the voxel set),
 *       - voxel id = rank of the voxel's slot among the occupied slots of the final table,
 *       - sums accumulate in point order; accum_double=0: float like the reference's atomics, =1: double
 *         accumulators rounded once at the end (what the CUDA path does).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int num_buckets, num_voxels, max_scan;
  float res;
  int* bucket_coord; /* B*3, (0,0,0) when empty  (voxel_coord_select_kernel :61-73) */
  int* bucket_id;    /* B, -1 when empty */
  int* vox_n;        /* V */
  float* vox_mean;   /* V*3 */
  float* vox_cov;    /* V*9 column-major */
} OrcVoxelMap;

static int coord_less(const int* a, const int* b) {
  if (a[0] != b[0]) return a[0] < b[0];
  if (a[1] != b[1]) return a[1] < b[1];
  return a[2] < b[2];
}
static int cmp_coord(const void* a, const void* b) {
  const int* x = (const int*)a; const int* y = (const int*)b;
  if (coord_less(x, y)) return -1;
  if (coord_less(y, x)) return 1;
  return 0;
}

ORC_API void orc_voxelmap_free(OrcVoxelMap* m) {
  if (!m) return;
  free(m->bucket_coord); free(m->bucket_id); free(m->vox_n); free(m->vox_mean); free(m->vox_cov);
  free(m);
}

/* lookup: find_voxel_correspondences.cu:32-60 -- stop at first empty bucket, at most max_scan probes */
static inline int voxelmap_lookup(const OrcVoxelMap* m, const int* c) {
  uint64_t h = orc_vector3i_hash(c[0], c[1], c[2]);
  for (int i = 0; i < m->max_scan; i++) {
    uint64_t b = (h + (uint64_t)i) % (uint64_t)m->num_buckets;
    if (m->bucket_id[b] < 0) return -1;
    const int* bc = m->bucket_coord + 3 * b;
    if (bc[0] == c[0] && bc[1] == c[1] && bc[2] == c[2]) return m->bucket_id[b];
  }
  return -1;
}

static OrcVoxelMap* voxelmap_build_impl(const float* pts, const float* cov9, int n, float res, int init_buckets, int max_scan, int accum_double, int ndt);

ORC_API OrcVoxelMap* orc_voxelmap_build(const float* pts, const float* cov9, int n, float res, int init_buckets, int max_scan, int accum_double) {
  return voxelmap_build_impl(pts, cov9, n, res, init_buckets, max_scan, accum_double, 0);
}

/* NDT voxel map: GaussianVoxelMap::create_voxelmap(points) gaussian_voxelmap.cu:209-231 -- accumulate p and p p^T
 * (:122-148), ndt_finalize_voxels_kernel (:178-198): mean = sum/n, cov = (sum_ppT - mean * sum^T)/n, then
 * covariance_regularization(MIN_EIG) on the voxel covariances (ndt_cuda.cu:129,140). */
ORC_API OrcVoxelMap* orc_ndt_voxelmap_build(const float* pts, int n, float res, int init_buckets, int max_scan, int accum_double) {
  return voxelmap_build_impl(pts, NULL, n, res, init_buckets, max_scan, accum_double, 1);
}

static OrcVoxelMap* voxelmap_build_impl(const float* pts, const float* cov9, int n, float res, int init_buckets, int max_scan, int accum_double, int ndt) {
  OrcVoxelMap* m = (OrcVoxelMap*)calloc(1, sizeof(OrcVoxelMap));
  m->res = res;
  m->max_scan = max_scan;
  int* coords = (int*)malloc(sizeof(int) * 3 * (size_t)(n > 0 ? n : 1));
  orc_voxel_coords(pts, n, res, coords);
  /* distinct coords, lexicographic order */
  int* uniq = (int*)malloc(sizeof(int) * 3 * (size_t)(n > 0 ? n : 1));
  memcpy(uniq, coords, sizeof(int) * 3 * (size_t)n);
  qsort(uniq, (size_t)n, 3 * sizeof(int), cmp_coord);
  int nu = 0;
  for (int i = 0; i < n; i++)
    if (i == 0 || cmp_coord(uniq + 3 * i, uniq + 3 * (nu - 1)) != 0) { memmove(uniq + 3 * nu, uniq + 3 * i, 3 * sizeof(int)); nu++; }

  int* slot_owner = NULL; /* index into uniq or -1 */
  for (int B = init_buckets;; B *= 2) { /* :265 loop has no upper bound in the reference */
    free(slot_owner);
    slot_owner = (int*)malloc(sizeof(int) * (size_t)B);
    for (int b = 0; b < B; b++) slot_owner[b] = -1;
    for (int u = 0; u < nu; u++) { /* voxel_bucket_assignment_kernel :21-58, serial */
      const int* c = uniq + 3 * u;
      uint64_t h = orc_vector3i_hash(c[0], c[1], c[2]);
      for (int i = 0; i < max_scan; i++) {
        uint64_t b = (h + (uint64_t)i) % (uint64_t)B;
        if (slot_owner[b] < 0) { slot_owner[b] = u; break; }
      }
    }
    /* materialise table + ids */
    free(m->bucket_coord); free(m->bucket_id);
    m->num_buckets = B;
    m->bucket_coord = (int*)calloc((size_t)B * 3, sizeof(int));
    m->bucket_id = (int*)malloc(sizeof(int) * (size_t)B);
    int nv = 0;
    for (int b = 0; b < B; b++) {
      if (slot_owner[b] >= 0) {
        memcpy(m->bucket_coord + 3 * b, uniq + 3 * slot_owner[b], 3 * sizeof(int));
        m->bucket_id[b] = nv++;
      } else m->bucket_id[b] = -1;
    }
    m->num_voxels = nv;
    /* failures = points whose voxel is not in the table (:57, counted per point) */
    long fails = 0;
    for (int i = 0; i < n; i++) if (voxelmap_lookup(m, coords + 3 * i) < 0) fails++;
    if ((double)fails / (double)n < 0.01) break; /* :280 */
    if (B > (1 << 28)) break;
  }
  free(slot_owner);
  free(uniq);

  int V = m->num_voxels;
  m->vox_n = (int*)calloc((size_t)(V > 0 ? V : 1), sizeof(int));
  m->vox_mean = (float*)calloc((size_t)(V > 0 ? V : 1) * 3, sizeof(float));
  m->vox_cov = (float*)calloc((size_t)(V > 0 ? V : 1) * 9, sizeof(float));
  double* dsum = accum_double ? (double*)calloc((size_t)(V > 0 ? V : 1) * 12, sizeof(double)) : NULL;
  for (int i = 0; i < n; i++) { /* accumulate_points_kernel :76-120 */
    int id = voxelmap_lookup(m, coords + 3 * i);
    if (id < 0) continue;
    m->vox_n[id]++;
    float ppt[9];
    const float* add9 = cov9 ? cov9 + 9 * (size_t)i : ppt;
    if (ndt) { /* cov = mean * mean^T of the point, :139 */
      const float* q = pts + 3 * (size_t)i;
      for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) ppt[c * 3 + r] = q[r] * q[c];
    }
    if (accum_double) {
      for (int d = 0; d < 3; d++) dsum[(size_t)id * 12 + d] += (double)pts[3 * (size_t)i + d];
      if (ndt) { /* the CUDA path forms p p^T in double before adding */
        const float* q = pts + 3 * (size_t)i;
        for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) dsum[(size_t)id * 12 + 3 + c * 3 + r] += (double)q[r] * (double)q[c];
      } else {
        for (int d = 0; d < 9; d++) dsum[(size_t)id * 12 + 3 + d] += (double)add9[d];
      }
    } else {
      for (int d = 0; d < 3; d++) m->vox_mean[(size_t)id * 3 + d] += pts[3 * (size_t)i + d];
      for (int d = 0; d < 9; d++) m->vox_cov[(size_t)id * 9 + d] += add9[d];
    }
  }
  for (int v = 0; v < V; v++) { /* finalize_voxels_kernel :158-176 / ndt_finalize_voxels_kernel :178-198 */
    if (ndt) {
      if (accum_double) {
        double nn = (double)m->vox_n[v], mean[3];
        for (int d = 0; d < 3; d++) mean[d] = dsum[(size_t)v * 12 + d] / nn;
        for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++)
          m->vox_cov[(size_t)v * 9 + c * 3 + r] = (float)((dsum[(size_t)v * 12 + 3 + c * 3 + r] - mean[r] * dsum[(size_t)v * 12 + c]) / nn);
        for (int d = 0; d < 3; d++) m->vox_mean[(size_t)v * 3 + d] = (float)mean[d];
      } else {
        float nn = (float)m->vox_n[v], sum[3], mean[3];
        for (int d = 0; d < 3; d++) { sum[d] = m->vox_mean[(size_t)v * 3 + d]; mean[d] = sum[d] / nn; }
        for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) m->vox_cov[(size_t)v * 9 + c * 3 + r] = (m->vox_cov[(size_t)v * 9 + c * 3 + r] - mean[r] * sum[c]) / nn;
        for (int d = 0; d < 3; d++) m->vox_mean[(size_t)v * 3 + d] = mean[d];
      }
      continue;
    }
    if (accum_double) {
      double nn = (double)m->vox_n[v];
      for (int d = 0; d < 3; d++) m->vox_mean[(size_t)v * 3 + d] = (float)(dsum[(size_t)v * 12 + d] / nn);
      for (int d = 0; d < 9; d++) m->vox_cov[(size_t)v * 9 + d] = (float)(dsum[(size_t)v * 12 + 3 + d] / nn);
    } else {
      float nn = (float)m->vox_n[v];
      for (int d = 0; d < 3; d++) m->vox_mean[(size_t)v * 3 + d] /= nn;
      for (int d = 0; d < 9; d++) m->vox_cov[(size_t)v * 9 + d] /= nn;
    }
  }
  free(dsum);
  free(coords);
  if (ndt) orc_regularize(m->vox_cov, V, ORC_REG_MIN_EIG); /* ndt_cuda.cu:129,140 */
  return m;
}

ORC_API int orc_voxelmap_num_buckets(const OrcVoxelMap* m) { return m->num_buckets; }
ORC_API int orc_voxelmap_num_voxels(const OrcVoxelMap* m) { return m->num_voxels; }
ORC_API void orc_voxelmap_get(const OrcVoxelMap* m, int* bucket_coord, int* bucket_id, int* vox_n, float* vox_mean, float* vox_cov) {
  if (bucket_coord) memcpy(bucket_coord, m->bucket_coord, sizeof(int) * 3 * (size_t)m->num_buckets);
  if (bucket_id) memcpy(bucket_id, m->bucket_id, sizeof(int) * (size_t)m->num_buckets);
  if (vox_n) memcpy(vox_n, m->vox_n, sizeof(int) * (size_t)m->num_voxels);
  if (vox_mean) memcpy(vox_mean, m->vox_mean, sizeof(float) * 3 * (size_t)m->num_voxels);
  if (vox_cov) memcpy(vox_cov, m->vox_cov, sizeof(float) * 9 * (size_t)m->num_voxels);
}

/* ------------------------------------------------------------------------------------------------------------
 * (6) neighbour offsets  --  fast_vgicp_cuda.cu:42-95   (enum order gicp_settings.hpp:8: DIRECT27, DIRECT7, DIRECT1,
 *     DIRECT_RADIUS)
 * ---------------------------------------------------------------------------------------------------------- */
enum { ORC_DIRECT27 = 0, ORC_DIRECT7 = 1, ORC_DIRECT1 = 2, ORC_DIRECT_RADIUS = 3 };
ORC_API int orc_offsets(int method, double radius, int* out, int cap) {
  int n = 0;
#define PUSH(a, b, c) do { if (n < cap) { out[3 * n] = (a); out[3 * n + 1] = (b); out[3 * n + 2] = (c); } n++; } while (0)
  switch (method) {
    case ORC_DIRECT1: PUSH(0, 0, 0); break;
    case ORC_DIRECT7:
      PUSH(0, 0, 0); PUSH(1, 0, 0); PUSH(-1, 0, 0); PUSH(0, 1, 0); PUSH(0, -1, 0); PUSH(0, 0, 1); PUSH(0, 0, -1);
      break;
    case ORC_DIRECT27:
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) PUSH(i - 1, j - 1, k - 1);
      break;
    case ORC_DIRECT_RADIUS: {
      int range = (int)ceil(radius);
      for (int i = -range; i <= range; i++) for (int j = -range; j <= range; j++) for (int k = -range; k <= range; k++) {
        double nrm = sqrt((double)i * i + (double)j * j + (double)k * k);
        if (nrm <= radius + 1e-3) PUSH(i, j, k);
      }
    } break;
    default: return -1;
  }
#undef PUSH
  return n;
}

/* ------------------------------------------------------------------------------------------------------------
 * (7) voxel correspondences  --  find_voxel_correspondences.cu:32-60,83-111 ; pose cast fast_vgicp_cuda.cu:265-274
 *     p' = R*a + t in float: ((R_r0*a0 + R_r1*a1) + R_r2*a2) + t_r with the adds fused the way nvcc contracts
 *     them: fmaf(R_r2,a2, fmaf(R_r1,a1, R_r0*a0)) + t_r.  Output order offset-major, point-minor (:93-110).
 *     T is a 4x4 column-major float matrix (Eigen::Isometry3f::data()).
 * ---------------------------------------------------------------------------------------------------------- */
static inline void transform_pt(const float* T, const float* a, float* o) {
  for (int r = 0; r < 3; r++) o[r] = fmaf(T[8 + r], a[2], fmaf(T[4 + r], a[1], T[r] * a[0])) + T[12 + r];
}
ORC_API void orc_transform_points(const float* T, const float* pts, int n, float* out) {
  for (int i = 0; i < n; i++) transform_pt(T, pts + 3 * i, out + 3 * i);
}

ORC_API long orc_find_correspondences(const OrcVoxelMap* m, const float* src, int n, const float* Tlin, const int* offsets, int n_off, int* pairs, long cap) {
  long cnt = 0;
  for (int o = 0; o < n_off; o++)
    for (int i = 0; i < n; i++) {
      float p[3];
      int c[3];
      transform_pt(Tlin, src + 3 * i, p);
      orc_voxel_coord(p, m->res, c);
      c[0] += offsets[3 * o]; c[1] += offsets[3 * o + 1]; c[2] += offsets[3 * o + 2];
      int id = voxelmap_lookup(m, c);
      if (id < 0) continue;
      if (cnt < cap) { pairs[2 * cnt] = i; pairs[2 * cnt + 1] = id; }
      cnt++;
    }
  return cnt;
}

/* ------------------------------------------------------------------------------------------------------------
 * (8) derivatives  --  compute_derivatives.cu:50-92 (H,b,err), :105-135 (err only), :151-184 (reduction)
 *     per-correspondence terms in float exactly as the reference forms them; the sum over correspondences is taken
 *     in double (sum_float=0) or in float in list order (sum_float=1) -- the reference's thrust reduction tree is
 *     unspecified, the two bracket it.
 *     H36 column-major 6x6, b6; returns the error.
 * ---------------------------------------------------------------------------------------------------------- */
static double compute_derivatives_impl(const OrcVoxelMap* m, const float* src, const float* src_cov9, const int* pairs, long n_pairs, const float* Tlin, const float* Teval,
                                       double* H36, double* b6, int sum_float, int ndt);

ORC_API double orc_compute_derivatives(const OrcVoxelMap* m, const float* src, const float* src_cov9, const int* pairs, long n_pairs, const float* Tlin, const float* Teval,
                                       double* H36, double* b6, int sum_float) {
  return compute_derivatives_impl(m, src, src_cov9, pairs, n_pairs, Tlin, Teval, H36, b6, sum_float, 0);
}

/* NDT: ndt_compute_derivatives.cu:33-175.  src_cov9 == NULL: P2D (M = cov_B^-1, :73); else D2D (M = (cov_B + R C_A R^T)^-1, :146).
 * weight = cauchy(resolution, |e|) (:15-18,78,150); voxels with <= 6 points contribute nothing (:61,132), also in the
 * error-only call (the reference evaluates the same functor and drops H, b). */
ORC_API double orc_ndt_compute_derivatives(const OrcVoxelMap* m, const float* src, const float* src_cov9, const int* pairs, long n_pairs, const float* Tlin, const float* Teval,
                                           double* H36, double* b6, int sum_float) {
  return compute_derivatives_impl(m, src, src_cov9, pairs, n_pairs, Tlin, Teval, H36, b6, sum_float, 1);
}

static double compute_derivatives_impl(const OrcVoxelMap* m, const float* src, const float* src_cov9, const int* pairs, long n_pairs, const float* Tlin, const float* Teval,
                                       double* H36, double* b6, int sum_float, int ndt) {
  double Hd[36] = {0}, bd[6] = {0}, ed = 0.0;
  float Hf[36] = {0}, bf[6] = {0}, ef = 0.0f;
  int want = (H36 != NULL && b6 != NULL);
  float Rl[9], Rlt[9];
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) { M3(Rl, r, c) = Tlin[c * 4 + r]; M3(Rlt, c, r) = Tlin[c * 4 + r]; }
  for (long ci = 0; ci < n_pairs; ci++) {
    int ia = pairs[2 * ci], iv = pairs[2 * ci + 1];
    if (iv < 0) continue;
    int np = m->vox_n[iv];
    if (!ndt && want && np <= 0) continue; /* :62-64 (the error-only functor has no such guard, :105-135) */
    if (ndt && np <= 6) continue;
    const float* meanA = src + 3 * (size_t)ia;
    static const float zero9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const float* covA = src_cov9 ? src_cov9 + 9 * (size_t)ia : zero9;
    const float* meanB = m->vox_mean + 3 * (size_t)iv;
    const float* covB = m->vox_cov + 9 * (size_t)iv;
    float pa[3];
    transform_pt(Teval, meanA, pa);
    float tmp[9], RCR[9], S[9], Minv[9];
    mul3f(Rl, covA, tmp);
    mul3f(tmp, Rlt, RCR);
    for (int j = 0; j < 9; j++) S[j] = covB[j] + RCR[j];
    inv3f(S, Minv);
    float e[3] = {meanB[0] - pa[0], meanB[1] - pa[1], meanB[2] - pa[2]};
    float w;
    if (ndt) {
      float x = sqrtf((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]); /* error.norm() */
      float k_sq = m->res * m->res;
      w = k_sq / (k_sq + x * x);
    } else {
      w = sqrtf((float)np);
    }
    float Me[3];
    for (int r = 0; r < 3; r++) Me[r] = (M3(Minv, r, 0) * e[0] + M3(Minv, r, 1) * e[1]) + M3(Minv, r, 2) * e[2];
    float err = w * ((e[0] * Me[0] + e[1] * Me[1]) + e[2] * Me[2]);
    if (sum_float) ef += err; else ed += (double)err;
    if (!want) continue;
    /* J = [skew(pa) | -I]  (3x6), :83-87 */
    float J[18]; /* column-major 3x6 */
    memset(J, 0, sizeof(J));
    J[3 * 1 + 0] = -pa[2]; J[3 * 2 + 0] = pa[1];
    J[3 * 0 + 1] = pa[2];  J[3 * 2 + 1] = -pa[0];
    J[3 * 0 + 2] = -pa[1]; J[3 * 1 + 2] = pa[0];
    J[3 * 3 + 0] = -1.0f; J[3 * 4 + 1] = -1.0f; J[3 * 5 + 2] = -1.0f;
    float JtM[18]; /* 6x3 : (w*J^T)*M , stored [row*3+col] */
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 3; c++) JtM[r * 3 + c] = ((w * J[3 * r + 0]) * M3(Minv, 0, c) + (w * J[3 * r + 1]) * M3(Minv, 1, c)) + (w * J[3 * r + 2]) * M3(Minv, 2, c);
    for (int r = 0; r < 6; r++) {
      for (int c = 0; c < 6; c++) {
        float h = (JtM[r * 3 + 0] * J[3 * c + 0] + JtM[r * 3 + 1] * J[3 * c + 1]) + JtM[r * 3 + 2] * J[3 * c + 2];
        if (sum_float) Hf[c * 6 + r] += h; else Hd[c * 6 + r] += (double)h;
      }
      float bb = (JtM[r * 3 + 0] * e[0] + JtM[r * 3 + 1] * e[1]) + JtM[r * 3 + 2] * e[2];
      if (sum_float) bf[r] += bb; else bd[r] += (double)bb;
    }
  }
  if (want) {
    for (int j = 0; j < 36; j++) H36[j] = sum_float ? (double)Hf[j] : Hd[j];
    for (int j = 0; j < 6; j++) b6[j] = sum_float ? (double)bf[j] : bd[j];
  }
  return sum_float ? (double)ef : ed;
}

/* ------------------------------------------------------------------------------------------------------------
 * (9) SE(3) exp + LM/GN optimizer (host, double)  --  so3.hpp:58-104 ; lsq_registration_impl.hpp:9-22,53-168
 *     Poses: 4x4 column-major double (Eigen::Isometry3d::data()).
 * ---------------------------------------------------------------------------------------------------------- */
ORC_API void orc_se3_exp(const double* a, double* T) {
  double om[3] = {a[0], a[1], a[2]};
  double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double theta = sqrt(theta_sq);
  double imag, real; /* so3_exp :58-77 */
  if (theta_sq < 1e-10) {
    double tq = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * tq;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * tq;
  } else {
    double th = sqrt(theta_sq), half = 0.5 * th;
    imag = sin(half) / th;
    real = cos(half);
  }
  double qw = real, qx = imag * om[0], qy = imag * om[1], qz = imag * om[2];
  /* Eigen QuaternionBase::toRotationMatrix */
  double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  double R[9]; /* column-major */
  M3(R, 0, 0) = 1 - (tyy + tzz); M3(R, 0, 1) = txy - twz; M3(R, 0, 2) = txz + twy;
  M3(R, 1, 0) = txy + twz; M3(R, 1, 1) = 1 - (txx + tzz); M3(R, 1, 2) = tyz - twx;
  M3(R, 2, 0) = txz - twy; M3(R, 2, 1) = tyz + twx; M3(R, 2, 2) = 1 - (txx + tyy);
  double Om[9] = {0, om[2], -om[1], -om[2], 0, om[0], om[1], -om[0], 0}; /* skewd, column-major */
  double Om2[9];
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) M3(Om2, r, c) = M3(Om, r, 0) * M3(Om, 0, c) + M3(Om, r, 1) * M3(Om, 1, c) + M3(Om, r, 2) * M3(Om, 2, c);
  double V[9];
  if (theta < 1e-10) {
    memcpy(V, R, sizeof(V)); /* so3.matrix() */
  } else {
    double tsq = theta * theta;
    double c1 = (1.0 - cos(theta)) / tsq, c2 = (theta - sin(theta)) / (tsq * theta);
    for (int j = 0; j < 9; j++) V[j] = ((j % 4 == 0) ? 1.0 : 0.0) + c1 * Om[j] + c2 * Om2[j];
  }
  memset(T, 0, 16 * sizeof(double));
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) T[c * 4 + r] = M3(R, r, c);
  for (int r = 0; r < 3; r++) T[12 + r] = M3(V, r, 0) * a[3] + M3(V, r, 1) * a[4] + M3(V, r, 2) * a[5];
  T[15] = 1.0;
}

static void iso_mul(const double* A, const double* B, double* out) { /* out = A*B (4x4 col-major affine) */
  double t[16];
  for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) {
    double s = 0;
    for (int k = 0; k < 4; k++) s += A[k * 4 + r] * B[c * 4 + k];
    t[c * 4 + r] = s;
  }
  memcpy(out, t, sizeof(t));
}

/* Eigen::LDLT<Matrix<double,6,6>> solve restated: symmetric pivoting on the largest remaining diagonal */
ORC_API void orc_ldlt_solve6(const double* A_in, const double* rhs, double* x) {
  double A[36];
  memcpy(A, A_in, sizeof(A));
  int perm[6];
  for (int i = 0; i < 6; i++) perm[i] = i;
#define A6(r, c) A[(c)*6 + (r)]
  for (int k = 0; k < 6; k++) {
    int p = k;
    double best = fabs(A6(k, k));
    for (int i = k + 1; i < 6; i++) if (fabs(A6(i, i)) > best) { best = fabs(A6(i, i)); p = i; }
    if (p != k) {
      for (int j = 0; j < 6; j++) { double t = A6(k, j); A6(k, j) = A6(p, j); A6(p, j) = t; }
      for (int j = 0; j < 6; j++) { double t = A6(j, k); A6(j, k) = A6(j, p); A6(j, p) = t; }
      int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    double d = A6(k, k);
    if (d == 0.0) continue;
    for (int i = k + 1; i < 6; i++) A6(i, k) /= d;
    for (int j = k + 1; j < 6; j++) for (int i = j; i < 6; i++) { A6(i, j) -= A6(i, k) * d * A6(j, k); A6(j, i) = A6(i, j); }
  }
  double y[6];
  for (int i = 0; i < 6; i++) y[i] = rhs[perm[i]];
  for (int i = 0; i < 6; i++) for (int j = 0; j < i; j++) y[i] -= A6(i, j) * y[j];
  for (int i = 0; i < 6; i++) y[i] = (A6(i, i) != 0.0) ? y[i] / A6(i, i) : 0.0;
  for (int i = 5; i >= 0; i--) for (int j = i + 1; j < 6; j++) y[i] -= A6(j, i) * y[j];
  for (int i = 0; i < 6; i++) x[perm[i]] = y[i];
#undef A6
}

typedef struct {
  int max_iterations;            /* 64   lsq_registration_impl.hpp:12 */
  double rotation_epsilon;       /* 2e-3 :13 */
  double transformation_epsilon; /* 5e-4 :14 */
  int use_gauss_newton;          /* 0 = LevenbergMarquardt :16 */
  int lm_max_iterations;         /* 10   :18 */
  double lm_init_lambda_factor;  /* 1e-9 :19 */
} OrcLsqParams;

ORC_API void orc_lsq_default_params(OrcLsqParams* p) {
  p->max_iterations = 64; p->rotation_epsilon = 2e-3; p->transformation_epsilon = 5e-4;
  p->use_gauss_newton = 0; p->lm_max_iterations = 10; p->lm_init_lambda_factor = 1e-9;
}

static int is_converged(const OrcLsqParams* p, const double* delta) { /* :82-91 */
  double mr = 0, mt = 0;
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) {
    double v = fabs(delta[c * 4 + r] - (r == c ? 1.0 : 0.0)) / p->rotation_epsilon;
    if (v > mr) mr = v;
  }
  for (int r = 0; r < 3; r++) { double v = fabs(delta[12 + r]) / p->transformation_epsilon; if (v > mt) mt = v; }
  return (mr > mt ? mr : mt) < 1.0;
}

/* problem callbacks: linearize(T,H,b) -> y ; error(T) -> y */
typedef double (*orc_linearize_fn)(void* ctx, const double* T, double* H36, double* b6);
typedef double (*orc_error_fn)(void* ctx, const double* T);

typedef struct {
  double T[16];
  double H[36];
  int iterations; /* nr_iterations_ */
  int converged;
  int n_linearize, n_error;
} OrcLsqResult;

static void lsq_optimize(const OrcLsqParams* p, void* ctx, orc_linearize_fn lin, orc_error_fn errf, const double* guess, OrcLsqResult* res) {
  double x0[16];
  memcpy(x0, guess, sizeof(x0));
  double lambda = -1.0;
  int converged = 0;
  memset(res->H, 0, sizeof(res->H));
  for (int i = 0; i < 6; i++) res->H[i * 7] = 1.0; /* final_hessian_.setIdentity() :21 */
  res->iterations = 0; res->n_linearize = res->n_error = 0;
  for (int it = 0; it < p->max_iterations && !converged; it++) { /* :65 */
    res->iterations = it;
    double H[36], b[6], delta[16], d[6], nb[6];
    double y0 = lin(ctx, x0, H, b);
    res->n_linearize++;
    for (int j = 0; j < 6; j++) nb[j] = -b[j];
    int ok = 0;
    if (p->use_gauss_newton) { /* step_gn :106-120 */
      orc_ldlt_solve6(H, nb, d);
      orc_se3_exp(d, delta);
      iso_mul(delta, x0, x0);
      memcpy(res->H, H, sizeof(H));
      ok = 1;
    } else { /* step_lm :123-168 */
      if (lambda < 0.0) {
        double mx = 0;
        for (int j = 0; j < 6; j++) if (fabs(H[j * 7]) > mx) mx = fabs(H[j * 7]);
        lambda = p->lm_init_lambda_factor * mx;
      }
      double nu = 2.0;
      for (int j = 0; j < p->lm_max_iterations; j++) {
        double Hl[36];
        memcpy(Hl, H, sizeof(H));
        for (int q = 0; q < 6; q++) Hl[q * 7] += lambda;
        orc_ldlt_solve6(Hl, nb, d);
        orc_se3_exp(d, delta);
        double xi[16];
        iso_mul(delta, x0, xi);
        double yi = errf(ctx, xi);
        res->n_error++;
        double den = 0;
        for (int q = 0; q < 6; q++) den += d[q] * (lambda * d[q] - b[q]);
        double rho = (y0 - yi) / den;
        if (rho < 0) {
          if (is_converged(p, delta)) { ok = 1; break; }
          lambda = nu * lambda;
          nu = 2 * nu;
          continue;
        }
        memcpy(x0, xi, sizeof(xi));
        double f = 1 - pow(2 * rho - 1, 3);
        lambda = lambda * (f > 1.0 / 3.0 ? f : 1.0 / 3.0);
        memcpy(res->H, H, sizeof(H));
        ok = 1;
        break;
      }
    }
    if (!ok) break; /* "lm not converged!!" :69-72 */
    converged = is_converged(p, delta);
  }
  memcpy(res->T, x0, sizeof(x0));
  res->converged = converged;
}

/* --- the float (CUDA-path) problem: FastVGICPCuda::linearize / compute_error, fast_vgicp_cuda_impl.hpp:170-178 --- */
typedef struct {
  const OrcVoxelMap* map;
  const float* src;
  const float* src_cov9;
  int n_src;
  const int* offsets;
  int n_off;
  int sum_float;
  int ndt;
  float Tlin[16];
  int* pairs;
  long n_pairs, cap_pairs;
} F32Problem;

static void to_f32_pose(const double* T, float* Tf) { for (int i = 0; i < 16; i++) Tf[i] = (float)T[i]; }

static double f32_linearize(void* c, const double* T, double* H, double* b) {
  F32Problem* p = (F32Problem*)c;
  to_f32_pose(T, p->Tlin); /* update_correspondences: fast_vgicp_cuda.cu:265-274 */
  p->n_pairs = orc_find_correspondences(p->map, p->src, p->n_src, p->Tlin, p->offsets, p->n_off, p->pairs, p->cap_pairs);
  float Te[16];
  to_f32_pose(T, Te);
  return compute_derivatives_impl(p->map, p->src, p->src_cov9, p->pairs, p->n_pairs, p->Tlin, Te, H, b, p->sum_float, p->ndt);
}
static double f32_error(void* c, const double* T) {
  F32Problem* p = (F32Problem*)c;
  float Te[16];
  to_f32_pose(T, Te);
  return compute_derivatives_impl(p->map, p->src, p->src_cov9, p->pairs, p->n_pairs, p->Tlin, Te, NULL, NULL, p->sum_float, p->ndt);
}

/* align(): LsqRegistration::computeTransformation with the float problem; guess/T 4x4 col-major double */
ORC_API int orc_align_f32(const OrcVoxelMap* map, const float* src, const float* src_cov9, int n_src, const int* offsets, int n_off, const OrcLsqParams* params,
                          const double* guess, int sum_float, OrcLsqResult* res) {
  F32Problem p;
  memset(&p, 0, sizeof(p));
  p.map = map; p.src = src; p.src_cov9 = src_cov9; p.n_src = n_src; p.offsets = offsets; p.n_off = n_off; p.sum_float = sum_float;
  p.cap_pairs = (long)n_src * n_off;
  p.pairs = (int*)malloc(sizeof(int) * 2 * (size_t)(p.cap_pairs > 0 ? p.cap_pairs : 1));
  lsq_optimize(params, &p, f32_linearize, f32_error, guess, res);
  free(p.pairs);
  return 0;
}

/* NDTCuda::align (ndt_cuda_impl.hpp:70-90): src = source points (P2D, src_cov9 NULL) or source voxel means + covariances (D2D) */
ORC_API int orc_align_ndt(const OrcVoxelMap* map, const float* src, const float* src_cov9, int n_src, const int* offsets, int n_off, const OrcLsqParams* params,
                          const double* guess, int sum_float, OrcLsqResult* res) {
  F32Problem p;
  memset(&p, 0, sizeof(p));
  p.map = map; p.src = src; p.src_cov9 = src_cov9; p.n_src = n_src; p.offsets = offsets; p.n_off = n_off; p.sum_float = sum_float; p.ndt = 1;
  p.cap_pairs = (long)n_src * n_off;
  p.pairs = (int*)malloc(sizeof(int) * 2 * (size_t)(p.cap_pairs > 0 ? p.cap_pairs : 1));
  lsq_optimize(params, &p, f32_linearize, f32_error, guess, res);
  free(p.pairs);
  return 0;
}

ORC_API double orc_evaluate_ndt(const OrcVoxelMap* map, const float* src, const float* src_cov9, int n_src, const int* offsets, int n_off, const double* Tlin_d,
                                const double* Teval_d, double* H36, double* b6, int sum_float, long* n_corr) {
  float Tl[16], Te[16];
  to_f32_pose(Tlin_d, Tl);
  to_f32_pose(Teval_d, Te);
  long cap = (long)n_src * n_off;
  int* pairs = (int*)malloc(sizeof(int) * 2 * (size_t)(cap > 0 ? cap : 1));
  long np = orc_find_correspondences(map, src, n_src, Tl, offsets, n_off, pairs, cap);
  double e = orc_ndt_compute_derivatives(map, src, src_cov9, pairs, np, Tl, Te, H36, b6, sum_float);
  if (n_corr) *n_corr = np;
  free(pairs);
  return e;
}

/* one evaluation, for stage-level tests: update_correspondences(Tlin) + compute_error(Teval,H,b) */
ORC_API double orc_evaluate_f32(const OrcVoxelMap* map, const float* src, const float* src_cov9, int n_src, const int* offsets, int n_off, const double* Tlin_d,
                                const double* Teval_d, double* H36, double* b6, int sum_float, long* n_corr) {
  float Tl[16], Te[16];
  to_f32_pose(Tlin_d, Tl);
  to_f32_pose(Teval_d, Te);
  long cap = (long)n_src * n_off;
  int* pairs = (int*)malloc(sizeof(int) * 2 * (size_t)(cap > 0 ? cap : 1));
  long np = orc_find_correspondences(map, src, n_src, Tl, offsets, n_off, pairs, cap);
  double e = orc_compute_derivatives(map, src, src_cov9, pairs, np, Tl, Te, H36, b6, sum_float);
  if (n_corr) *n_corr = np;
  free(pairs);
  return e;
}

/* ============================================================================================================
 * (10) CPU twin in double: FastVGICP (OpenMP) -- the reference's own CPU implementation of this path, restated
 *      for the cpu_baseline / --impl reference timing and as a second opinion on the converged pose.
 *      fast_gicp_impl.hpp:243-301 (covariances), fast_vgicp_voxel.hpp:129-182 (voxel map, unordered_map, double
 *      coords), fast_vgicp_impl.hpp:73-204 (correspondences + Mahalanobis precompute + linearize/compute_error).
 * ========================================================================================================== */
static void jacobi_eig3d(const double* A_in, double* evals, double* V) { /* symmetric 3x3, cyclic Jacobi; stands in for JacobiSVD */
  double A[9];
  memcpy(A, A_in, sizeof(A));
  for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 32; sweep++) {
    double off = fabs(M3(A, 0, 1)) + fabs(M3(A, 0, 2)) + fabs(M3(A, 1, 2));
    if (off < 1e-300) break;
    for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
      double apq = M3(A, p, q);
      if (fabs(apq) < 1e-300) continue;
      double th = (M3(A, q, q) - M3(A, p, p)) / (2.0 * apq);
      double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; k++) { double akp = M3(A, k, p), akq = M3(A, k, q); M3(A, k, p) = c * akp - s * akq; M3(A, k, q) = s * akp + c * akq; }
      for (int k = 0; k < 3; k++) { double apk = M3(A, p, k), aqk = M3(A, q, k); M3(A, p, k) = c * apk - s * aqk; M3(A, q, k) = s * apk + c * aqk; }
      for (int k = 0; k < 3; k++) { double vkp = M3(V, k, p), vkq = M3(V, k, q); M3(V, k, p) = c * vkp - s * vkq; M3(V, k, q) = s * vkp + c * vkq; }
    }
  }
  evals[0] = M3(A, 0, 0); evals[1] = M3(A, 1, 1); evals[2] = M3(A, 2, 2);
}

static void inv3d(const double* m, double* out) {
  double c00 = M3(m, 1, 1) * M3(m, 2, 2) - M3(m, 1, 2) * M3(m, 2, 1);
  double c10 = M3(m, 2, 1) * M3(m, 0, 2) - M3(m, 2, 2) * M3(m, 0, 1);
  double c20 = M3(m, 0, 1) * M3(m, 1, 2) - M3(m, 0, 2) * M3(m, 1, 1);
  double det = c00 * M3(m, 0, 0) + c10 * M3(m, 1, 0) + c20 * M3(m, 2, 0);
  double id = 1.0 / det;
  M3(out, 0, 0) = c00 * id; M3(out, 0, 1) = c10 * id; M3(out, 0, 2) = c20 * id;
  M3(out, 1, 0) = (M3(m, 1, 2) * M3(m, 2, 0) - M3(m, 1, 0) * M3(m, 2, 2)) * id;
  M3(out, 1, 1) = (M3(m, 2, 2) * M3(m, 0, 0) - M3(m, 2, 0) * M3(m, 0, 2)) * id;
  M3(out, 1, 2) = (M3(m, 0, 2) * M3(m, 1, 0) - M3(m, 0, 0) * M3(m, 1, 2)) * id;
  M3(out, 2, 0) = (M3(m, 1, 0) * M3(m, 2, 1) - M3(m, 1, 1) * M3(m, 2, 0)) * id;
  M3(out, 2, 1) = (M3(m, 2, 0) * M3(m, 0, 1) - M3(m, 2, 1) * M3(m, 0, 0)) * id;
  M3(out, 2, 2) = (M3(m, 0, 0) * M3(m, 1, 1) - M3(m, 0, 1) * M3(m, 1, 0)) * id;
}

/* FastGICP::calculate_covariances (fast_gicp_impl.hpp:243-301): kd-tree kNN, centred double covariance, PLANE via SVD
 * (values (1,1,1e-3) on singular values descending == 1e-3 on the smallest eigen direction). cov_out: N x 9 double */
ORC_API int orc64_covariances(const float* pts, int n, int k, int reg_method, double* cov_out, int num_threads) {
  if (k > n || k > 256) return -1;
  KdTree* t = kd_build(pts, n);
#ifdef _OPENMP
  if (num_threads <= 0) num_threads = omp_get_max_threads();
#else
  num_threads = 1;
#endif
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
  for (int i = 0; i < n; i++) {
    float bd[256];
    int bi[256];
    int cnt = 0;
    kd_search(t, 0, pts + 3 * i, k, bd, bi, &cnt);
    double mean[3] = {0, 0, 0};
    for (int j = 0; j < k; j++) for (int d = 0; d < 3; d++) mean[d] += (double)pts[3 * (size_t)bi[j] + d];
    for (int d = 0; d < 3; d++) mean[d] /= k;
    double C[9] = {0};
    for (int j = 0; j < k; j++) {
      double q[3];
      for (int d = 0; d < 3; d++) q[d] = (double)pts[3 * (size_t)bi[j] + d] - mean[d];
      for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) M3(C, r, c) += q[r] * q[c];
    }
    for (int j = 0; j < 9; j++) C[j] /= k;
    double* out = cov_out + 9 * (size_t)i;
    if (reg_method == ORC_REG_NONE) {
      memcpy(out, C, sizeof(C));
    } else if (reg_method == ORC_REG_FROBENIUS) {
      double Cl[9], Ci[9];
      memcpy(Cl, C, sizeof(C));
      Cl[0] += 1e-3; Cl[4] += 1e-3; Cl[8] += 1e-3;
      inv3d(Cl, Ci);
      double nn = 0;
      for (int j = 0; j < 9; j++) nn += Ci[j] * Ci[j];
      nn = sqrt(nn);
      for (int j = 0; j < 9; j++) Ci[j] /= nn;
      inv3d(Ci, out);
    } else {
      double ev[3], V[9];
      jacobi_eig3d(C, ev, V);
      /* sort descending like singular values */
      int o[3] = {0, 1, 2};
      for (int a = 0; a < 2; a++) for (int b2 = a + 1; b2 < 3; b2++) if (ev[o[b2]] > ev[o[a]]) { int tt = o[a]; o[a] = o[b2]; o[b2] = tt; }
      double vals[3];
      if (reg_method == ORC_REG_PLANE) { vals[0] = 1; vals[1] = 1; vals[2] = 1e-3; }
      else if (reg_method == ORC_REG_MIN_EIG) { for (int a = 0; a < 3; a++) vals[a] = fmax(ev[o[a]], 1e-3); }
      else { double mx = ev[o[0]]; for (int a = 0; a < 3; a++) vals[a] = fmax(ev[o[a]] / mx, 1e-3); }
      for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int a = 0; a < 3; a++) s += M3(V, r, o[a]) * vals[a] * M3(V, c, o[a]);
        M3(out, r, c) = s;
      }
    }
  }
  kd_free(t);
  return 0;
}

/* CPU voxel map: unordered_map keyed by voxel coord (double arithmetic, fast_vgicp_voxel.hpp:158-160); here an
 * open-addressing table without probe cap -- the CPU map never drops voxels. Additive voxels (:105-122). */
typedef struct {
  int cap; /* pow2 */
  int* key; /* cap*3 */
  int* id;  /* cap, -1 empty */
  int nv;
  int* vn;
  double* vmean; /* nv*3 */
  double* vcov;  /* nv*9 */
  double res;
} Map64;

static int map64_find(const Map64* m, const int* c, int insert, Map64* mm) {
  uint64_t h = orc_vector3i_hash(c[0], c[1], c[2]);
  for (uint64_t i = 0;; i++) {
    uint64_t b = (h + i) & (uint64_t)(m->cap - 1);
    if (m->id[b] < 0) {
      if (!insert) return -1;
      mm->key[3 * b] = c[0]; mm->key[3 * b + 1] = c[1]; mm->key[3 * b + 2] = c[2];
      mm->id[b] = mm->nv++;
      return mm->id[b];
    }
    if (m->key[3 * b] == c[0] && m->key[3 * b + 1] == c[1] && m->key[3 * b + 2] == c[2]) return m->id[b];
  }
}
static inline void coord64(const double* p, double res, int* c) { for (int d = 0; d < 3; d++) c[d] = (int)floor(p[d] / res - 0.5); }

static Map64* map64_build(const float* pts, const double* cov9, int n, double res) {
  Map64* m = (Map64*)calloc(1, sizeof(Map64));
  int cap = 1024;
  while (cap < 4 * n) cap *= 2;
  m->cap = cap; m->res = res;
  m->key = (int*)calloc((size_t)cap * 3, sizeof(int));
  m->id = (int*)malloc(sizeof(int) * (size_t)cap);
  for (int i = 0; i < cap; i++) m->id[i] = -1;
  int* pid = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    double p[3] = {pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]};
    int c[3];
    coord64(p, res, c);
    pid[i] = map64_find(m, c, 1, m);
  }
  m->vn = (int*)calloc((size_t)(m->nv > 0 ? m->nv : 1), sizeof(int));
  m->vmean = (double*)calloc((size_t)(m->nv > 0 ? m->nv : 1) * 3, sizeof(double));
  m->vcov = (double*)calloc((size_t)(m->nv > 0 ? m->nv : 1) * 9, sizeof(double));
  for (int i = 0; i < n; i++) {
    int v = pid[i];
    m->vn[v]++;
    for (int d = 0; d < 3; d++) m->vmean[3 * (size_t)v + d] += (double)pts[3 * (size_t)i + d];
    for (int d = 0; d < 9; d++) m->vcov[9 * (size_t)v + d] += cov9[9 * (size_t)i + d];
  }
  for (int v = 0; v < m->nv; v++) {
    for (int d = 0; d < 3; d++) m->vmean[3 * (size_t)v + d] /= m->vn[v];
    for (int d = 0; d < 9; d++) m->vcov[9 * (size_t)v + d] /= m->vn[v];
  }
  free(pid);
  return m;
}
static void map64_free(Map64* m) {
  if (!m) return;
  free(m->key); free(m->id); free(m->vn); free(m->vmean); free(m->vcov); free(m);
}

typedef struct {
  const float* src; const double* src_cov; int n_src;
  const float* tgt; const double* tgt_cov; int n_tgt;
  double res; const int* offsets; int n_off; int num_threads;
  Map64* map;
  int* corr; /* pairs (src idx, voxel id) */
  double* mahal; /* 9 per corr */
  long n_corr, cap_corr;
} F64Problem;

static void f64_update_correspondences(F64Problem* p, const double* T) { /* fast_vgicp_impl.hpp:73-116 */
  /* parallel lookup like the reference (:82-97: per-thread lists, concatenated); two passes keep the list in point
   * order whatever the thread count */
  int* per_point = (int*)malloc(sizeof(int) * (size_t)(p->n_src + 1));
  int* found = (int*)malloc(sizeof(int) * (size_t)(p->n_src > 0 ? p->n_src : 1) * (size_t)p->n_off);
#pragma omp parallel for num_threads(p->num_threads) schedule(guided, 8)
  for (int i = 0; i < p->n_src; i++) {
    double a[3] = {p->src[3 * (size_t)i], p->src[3 * (size_t)i + 1], p->src[3 * (size_t)i + 2]}, q[3];
    for (int r = 0; r < 3; r++) q[r] = T[r] * a[0] + T[4 + r] * a[1] + T[8 + r] * a[2] + T[12 + r];
    int c[3];
    coord64(q, p->res, c);
    int m = 0;
    for (int o = 0; o < p->n_off; o++) {
      int cc[3] = {c[0] + p->offsets[3 * o], c[1] + p->offsets[3 * o + 1], c[2] + p->offsets[3 * o + 2]};
      int v = map64_find(p->map, cc, 0, NULL);
      if (v >= 0) found[(size_t)i * p->n_off + m++] = v;
    }
    per_point[i] = m;
  }
  long cnt = 0;
  for (int i = 0; i < p->n_src; i++) { int m = per_point[i]; per_point[i] = (int)cnt; cnt += m; }
  per_point[p->n_src] = (int)cnt;
#pragma omp parallel for num_threads(p->num_threads) schedule(static)
  for (int i = 0; i < p->n_src; i++)
    for (int j = per_point[i]; j < per_point[i + 1]; j++) { p->corr[2 * (size_t)j] = i; p->corr[2 * (size_t)j + 1] = found[(size_t)i * p->n_off + (j - per_point[i])]; }
  free(per_point);
  free(found);
  p->n_corr = cnt;
  double R[9], Rt[9];
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) { M3(R, r, c) = T[c * 4 + r]; M3(Rt, c, r) = T[c * 4 + r]; }
#pragma omp parallel for num_threads(p->num_threads) schedule(guided, 8)
  for (long ci = 0; ci < cnt; ci++) {
    const double* cA = p->src_cov + 9 * (size_t)p->corr[2 * ci];
    const double* cB = p->map->vcov + 9 * (size_t)p->corr[2 * ci + 1];
    double t1[9], S[9];
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) M3(t1, r, c) = M3(R, r, 0) * M3(cA, 0, c) + M3(R, r, 1) * M3(cA, 1, c) + M3(R, r, 2) * M3(cA, 2, c);
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) M3(S, r, c) = M3(cB, r, c) + (M3(t1, r, 0) * M3(Rt, 0, c) + M3(t1, r, 1) * M3(Rt, 1, c) + M3(t1, r, 2) * M3(Rt, 2, c));
    inv3d(S, p->mahal + 9 * (size_t)ci);
  }
}

static double f64_eval(F64Problem* p, const double* T, double* H, double* b) { /* :135-177, :182-204 */
  double sum = 0;
  int want = (H && b);
  int nt = p->num_threads;
  double* Hs = (double*)calloc((size_t)nt * 42, sizeof(double));
#pragma omp parallel for num_threads(nt) reduction(+ : sum) schedule(guided, 8)
  for (long ci = 0; ci < p->n_corr; ci++) {
    int ia = p->corr[2 * ci], iv = p->corr[2 * ci + 1];
    double a[3] = {p->src[3 * (size_t)ia], p->src[3 * (size_t)ia + 1], p->src[3 * (size_t)ia + 2]}, q[3], e[3], Me[3];
    for (int r = 0; r < 3; r++) q[r] = T[r] * a[0] + T[4 + r] * a[1] + T[8 + r] * a[2] + T[12 + r];
    const double* mB = p->map->vmean + 3 * (size_t)iv;
    const double* M = p->mahal + 9 * (size_t)ci;
    for (int r = 0; r < 3; r++) e[r] = mB[r] - q[r];
    for (int r = 0; r < 3; r++) Me[r] = M3(M, r, 0) * e[0] + M3(M, r, 1) * e[1] + M3(M, r, 2) * e[2];
    double w = sqrt((double)p->map->vn[iv]);
    sum += w * (e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2]);
    if (!want) continue;
    double J[18];
    memset(J, 0, sizeof(J));
    J[3 * 1 + 0] = -q[2]; J[3 * 2 + 0] = q[1]; J[3 * 0 + 1] = q[2]; J[3 * 2 + 1] = -q[0]; J[3 * 0 + 2] = -q[1]; J[3 * 1 + 2] = q[0];
    J[3 * 3 + 0] = -1; J[3 * 4 + 1] = -1; J[3 * 5 + 2] = -1;
    double JtM[18];
    for (int r = 0; r < 6; r++) for (int c = 0; c < 3; c++) JtM[r * 3 + c] = w * (J[3 * r] * M3(M, 0, c) + J[3 * r + 1] * M3(M, 1, c) + J[3 * r + 2] * M3(M, 2, c));
#ifdef _OPENMP
    double* acc = Hs + (size_t)omp_get_thread_num() * 42;
#else
    double* acc = Hs;
#endif
    for (int r = 0; r < 6; r++) {
      for (int c = 0; c < 6; c++) acc[c * 6 + r] += JtM[r * 3] * J[3 * c] + JtM[r * 3 + 1] * J[3 * c + 1] + JtM[r * 3 + 2] * J[3 * c + 2];
      acc[36 + r] += JtM[r * 3] * e[0] + JtM[r * 3 + 1] * e[1] + JtM[r * 3 + 2] * e[2];
    }
  }
  if (want) {
    memset(H, 0, 36 * sizeof(double));
    memset(b, 0, 6 * sizeof(double));
    for (int t = 0; t < nt; t++) { for (int j = 0; j < 36; j++) H[j] += Hs[(size_t)t * 42 + j]; for (int j = 0; j < 6; j++) b[j] += Hs[(size_t)t * 42 + 36 + j]; }
  }
  free(Hs);
  return sum;
}
static double f64_linearize(void* c, const double* T, double* H, double* b) {
  F64Problem* p = (F64Problem*)c;
  if (!p->map) p->map = map64_build(p->tgt, p->tgt_cov, p->n_tgt, p->res); /* lazy build :120-123 */
  f64_update_correspondences(p, T);
  return f64_eval(p, T, H, b);
}
static double f64_error(void* c, const double* T) { return f64_eval((F64Problem*)c, T, NULL, NULL); }

/* FastVGICP::align with precomputed double covariances (the "reuse" protocol keeps them; the full protocol calls
 * orc64_covariances for both clouds first). */
ORC_API int orc64_align(const float* tgt, const double* tgt_cov, int n_tgt, const float* src, const double* src_cov, int n_src, double res, const int* offsets, int n_off,
                        const OrcLsqParams* params, const double* guess, int num_threads, OrcLsqResult* res_out) {
  F64Problem p;
  memset(&p, 0, sizeof(p));
#ifdef _OPENMP
  if (num_threads <= 0) num_threads = omp_get_max_threads();
#else
  num_threads = 1;
#endif
  p.src = src; p.src_cov = src_cov; p.n_src = n_src; p.tgt = tgt; p.tgt_cov = tgt_cov; p.n_tgt = n_tgt;
  p.res = res; p.offsets = offsets; p.n_off = n_off; p.num_threads = num_threads;
  p.cap_corr = (long)n_src * n_off;
  p.corr = (int*)malloc(sizeof(int) * 2 * (size_t)(p.cap_corr > 0 ? p.cap_corr : 1));
  p.mahal = (double*)malloc(sizeof(double) * 9 * (size_t)(p.cap_corr > 0 ? p.cap_corr : 1));
  lsq_optimize(params, &p, f64_linearize, f64_error, guess, res_out);
  map64_free(p.map);
  free(p.corr); free(p.mahal);
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------------
 * FastGICP (BASELINE config 1: the CPU GICP row, single thread or OpenMP)  --  include/fast_gicp/gicp/impl/fast_gicp_impl.hpp
 *   update_correspondences :117-158 : nearest target point of every transformed source point (float pose and query, k = 1, kd-tree),
 *                                     accepted below corr_dist_threshold^2; mahalanobis = (C_B + T C_A T^T)^-1 (3x3 block)
 *   linearize :160-216, compute_error :218-240 : sum e^T M e ; H += J^T M J ; b += J^T M e with J = [skew(T a) | -I], double
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* src; const double* src_cov; int n_src;
  const float* tgt; const double* tgt_cov; int n_tgt;
  double corr_dist_sq; int num_threads;
  KdTree* tree;
  int* corr;      /* target index or -1 per source point */
  double* mahal;  /* 9 per source point */
} GicpProblem;

static void gicp_update_correspondences(GicpProblem* p, const double* T) {
  float Tf[16];
  for (int j = 0; j < 16; j++) Tf[j] = (float)T[j];
  double R[9], Rt[9];
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) { M3(R, r, c) = T[c * 4 + r]; M3(Rt, c, r) = T[c * 4 + r]; }
#pragma omp parallel for num_threads(p->num_threads) schedule(guided, 8)
  for (int i = 0; i < p->n_src; i++) {
    const float* a = p->src + 3 * (size_t)i;
    float q[3];
    for (int r = 0; r < 3; r++) q[r] = ((Tf[r] * a[0] + Tf[4 + r] * a[1]) + Tf[8 + r] * a[2]) + Tf[12 + r];
    float bd[1];
    int bi[1], cnt = 0;
    kd_search(p->tree, 0, q, 1, bd, bi, &cnt);
    p->corr[i] = (cnt > 0 && (double)bd[0] < p->corr_dist_sq) ? bi[0] : -1;
    if (p->corr[i] < 0) continue;
    const double* cA = p->src_cov + 9 * (size_t)i;
    const double* cB = p->tgt_cov + 9 * (size_t)p->corr[i];
    double t1[9], S[9];
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) M3(t1, r, c) = M3(R, r, 0) * M3(cA, 0, c) + M3(R, r, 1) * M3(cA, 1, c) + M3(R, r, 2) * M3(cA, 2, c);
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) M3(S, r, c) = M3(cB, r, c) + (M3(t1, r, 0) * M3(Rt, 0, c) + M3(t1, r, 1) * M3(Rt, 1, c) + M3(t1, r, 2) * M3(Rt, 2, c));
    inv3d(S, p->mahal + 9 * (size_t)i);
  }
}

static double gicp_eval(GicpProblem* p, const double* T, double* H, double* b) {
  double sum = 0;
  const int want = (H && b);
  const int nt = p->num_threads;
  double* Hs = (double*)calloc((size_t)nt * 42, sizeof(double));
#pragma omp parallel for num_threads(nt) reduction(+ : sum) schedule(guided, 8)
  for (int i = 0; i < p->n_src; i++) {
    const int it = p->corr[i];
    if (it < 0) continue;
    double a[3] = {p->src[3 * (size_t)i], p->src[3 * (size_t)i + 1], p->src[3 * (size_t)i + 2]}, q[3], e[3], Me[3];
    for (int r = 0; r < 3; r++) q[r] = T[r] * a[0] + T[4 + r] * a[1] + T[8 + r] * a[2] + T[12 + r];
    const double* M = p->mahal + 9 * (size_t)i;
    for (int r = 0; r < 3; r++) e[r] = (double)p->tgt[3 * (size_t)it + r] - q[r];
    for (int r = 0; r < 3; r++) Me[r] = M3(M, r, 0) * e[0] + M3(M, r, 1) * e[1] + M3(M, r, 2) * e[2];
    sum += e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2];
    if (!want) continue;
    double J[18];
    memset(J, 0, sizeof(J));
    J[3 * 1 + 0] = -q[2]; J[3 * 2 + 0] = q[1]; J[3 * 0 + 1] = q[2]; J[3 * 2 + 1] = -q[0]; J[3 * 0 + 2] = -q[1]; J[3 * 1 + 2] = q[0];
    J[3 * 3 + 0] = -1; J[3 * 4 + 1] = -1; J[3 * 5 + 2] = -1;
    double JtM[18];
    for (int r = 0; r < 6; r++) for (int c = 0; c < 3; c++) JtM[r * 3 + c] = J[3 * r] * M3(M, 0, c) + J[3 * r + 1] * M3(M, 1, c) + J[3 * r + 2] * M3(M, 2, c);
#ifdef _OPENMP
    double* acc = Hs + (size_t)omp_get_thread_num() * 42;
#else
    double* acc = Hs;
#endif
    for (int r = 0; r < 6; r++) {
      for (int c = 0; c < 6; c++) acc[c * 6 + r] += JtM[r * 3] * J[3 * c] + JtM[r * 3 + 1] * J[3 * c + 1] + JtM[r * 3 + 2] * J[3 * c + 2];
      acc[36 + r] += JtM[r * 3] * e[0] + JtM[r * 3 + 1] * e[1] + JtM[r * 3 + 2] * e[2];
    }
  }
  if (want) {
    memset(H, 0, 36 * sizeof(double));
    memset(b, 0, 6 * sizeof(double));
    for (int t = 0; t < nt; t++) { for (int j = 0; j < 36; j++) H[j] += Hs[(size_t)t * 42 + j]; for (int j = 0; j < 6; j++) b[j] += Hs[(size_t)t * 42 + 36 + j]; }
  }
  free(Hs);
  return sum;
}
static double gicp_linearize(void* c, const double* T, double* H, double* b) {
  gicp_update_correspondences((GicpProblem*)c, T);
  return gicp_eval((GicpProblem*)c, T, H, b);
}
static double gicp_error(void* c, const double* T) { return gicp_eval((GicpProblem*)c, T, NULL, NULL); }

/* FastGICP::align with precomputed double covariances (orc64_covariances); max_corr_dist <= 0: no threshold (the default,
 * corr_dist_threshold_ = float max, fast_gicp_impl.hpp:29). The kd-tree over the target is part of setInputTarget (:76-84). */
ORC_API int orc64_align_gicp(const float* tgt, const double* tgt_cov, int n_tgt, const float* src, const double* src_cov, int n_src, double max_corr_dist,
                             const OrcLsqParams* params, const double* guess, int num_threads, OrcLsqResult* res_out) {
  GicpProblem p;
  memset(&p, 0, sizeof(p));
#ifdef _OPENMP
  if (num_threads <= 0) num_threads = omp_get_max_threads();
#else
  num_threads = 1;
#endif
  p.src = src; p.src_cov = src_cov; p.n_src = n_src; p.tgt = tgt; p.tgt_cov = tgt_cov; p.n_tgt = n_tgt;
  p.corr_dist_sq = max_corr_dist > 0 ? max_corr_dist * max_corr_dist : 1e300;
  p.num_threads = num_threads;
  p.tree = kd_build(tgt, n_tgt);
  p.corr = (int*)malloc(sizeof(int) * (size_t)(n_src > 0 ? n_src : 1));
  p.mahal = (double*)malloc(sizeof(double) * 9 * (size_t)(n_src > 0 ? n_src : 1));
  lsq_optimize(params, &p, gicp_linearize, gicp_error, guess, res_out);
  kd_free(p.tree);
  free(p.corr); free(p.mahal);
  return 0;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Input preparation: what the reference's callers do to a raw scan before setInputTarget/Source.
 * ------------------------------------------------------------------------------------------------------------------- */

/* src/align.cpp:128-133 : erase(remove_if(squaredNorm() < 1e-3)) -- stable, float arithmetic ((x*x + y*y) + z*z as Eigen's
 * unrolled redux evaluates a 3-vector).  out may alias xyz.  Returns the number of points kept. */
ORC_API int orc_remove_near_origin(const float* xyz, int n, float* out) {
  int m = 0;
  for (int i = 0; i < n; i++) {
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const float sq = (x * x + y * y) + z * z;
    if (sq < 1e-3f) continue;
    out[3 * m] = x; out[3 * m + 1] = y; out[3 * m + 2] = z;
    m++;
  }
  return m;
}

/* pcl::ApproximateVoxelGrid<PointXYZ>::applyFilter (src/align.cpp:136-147, src/kitti.cpp:80-82, src/python/main.cpp:46-62,81-91).
 * PCL is not vendored: restated from the published algorithm (pcl/filters/impl/approximate_voxel_grid.hpp) -- a streaming
 * filter with a `histsize`-entry (default 512) hash history; a point goes to entry (ix*7171 + iy*3079 + iz*4231) & (histsize-1)
 * with ix = floor(x * inverse_leaf) in float; an entry holding a different voxel is flushed (centroid emitted) first; the
 * remaining entries are flushed in entry order at the end.  Centroids accumulate in float in input order.
 * Pinned by the point counts the reference prints (README.md:116: 17249 / 17518 for the data/ pair), tests/test_oracle_golden.py.
 * out: capacity n points.  Returns the number of output points. */
ORC_API int orc_approximate_voxel_grid(const float* xyz, int n, float leaf, int histsize, float* out) {
  typedef struct { int ix, iy, iz, count; float sx, sy, sz; } He;
  He* hist = (He*)calloc((size_t)histsize, sizeof(He));
  const float inv = 1.0f / leaf;
  int m = 0;
  for (int i = 0; i < n; i++) {
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const int ix = (int)floorf(x * inv), iy = (int)floorf(y * inv), iz = (int)floorf(z * inv);
    const long long hl = (long long)ix * 7171 + (long long)iy * 3079 + (long long)iz * 4231;
    He* h = &hist[(size_t)(hl & (long long)(histsize - 1))];
    if (h->count && (h->ix != ix || h->iy != iy || h->iz != iz)) {
      const float c = (float)h->count;
      out[3 * m] = h->sx / c; out[3 * m + 1] = h->sy / c; out[3 * m + 2] = h->sz / c;
      m++;
      h->count = 0; h->sx = h->sy = h->sz = 0.0f;
    }
    h->ix = ix; h->iy = iy; h->iz = iz;
    h->count++;
    h->sx += x; h->sy += y; h->sz += z;
  }
  for (int s = 0; s < histsize; s++) {
    He* h = &hist[s];
    if (!h->count) continue;
    const float c = (float)h->count;
    out[3 * m] = h->sx / c; out[3 * m + 1] = h->sy / c; out[3 * m + 2] = h->sz / c;
    m++;
  }
  free(hist);
  return m;
}
