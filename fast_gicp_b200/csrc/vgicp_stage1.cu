// vgicp_stage1.cu -- stage 1 kernels: exact k-NN, covariance estimation, regularisation, RBF covariances.
//
// This translation unit is compiled with --fmad=false: every fused multiply-add it wants is spelled __fmaf_rn, every
// other a*b+c stays two roundings, so the per-point results (neighbour order, raw covariance, eigen-decomposition,
// regularised covariance) are a pure function of the input bits and can be checked bit-for-bit against a CPU
// restatement compiled with -ffp-contract=off.  The closed-form 3x3 eigen-solver amplifies 1-ulp differences on
// line-like neighbourhoods (lidar rings) into percent-level differences of the covariance, so "close" is not
// good enough here.
#include "vgicp_stage1.cuh"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>

namespace vgicp {

// ---------------------------------------------------------------------------------------------------------------
// Stage 1a: exact brute-force k-NN inside one cloud (self included).  One query per thread, targets staged through
// shared memory in tiles (every thread reads the same target -> broadcast, conflict-free), per-thread ascending
// top-k list in shared memory laid out [k][thread] (conflict-free).  d2 = (dx*dx + dy*dy) + dz*dz, no contraction.
// Replaces brute_force_knn.cu:16-60 (global-memory heap per thread) and the CPU kd-tree of
// fast_vgicp_cuda_impl.hpp:152-167.  Output rows ascending in (d2, index).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kKnnThreads) k_knn_bruteforce(const float4* __restrict__ pts, int n, int k, int* __restrict__ nbr) {
  extern __shared__ float smem_knn[];
  float4* tile = reinterpret_cast<float4*>(smem_knn);                       // kKnnTile float4
  float* ld = smem_knn + 4 * kKnnTile;                                      // [k][kKnnThreads]
  int* li = reinterpret_cast<int*>(ld + (size_t)k * kKnnThreads);           // [k][kKnnThreads]
  const int tid = threadIdx.x;
  const int q = blockIdx.x * kKnnThreads + tid;
  const bool active = q < n;
  float4 qp = active ? pts[q] : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < k; j++) {
    ld[j * kKnnThreads + tid] = __int_as_float(0x7f800000);  // +inf
    li[j * kKnnThreads + tid] = -1;
  }
  float worst = __int_as_float(0x7f800000);
  for (int base = 0; base < n; base += kKnnTile) {
    __syncthreads();
    for (int j = tid; j < kKnnTile; j += kKnnThreads) {
      int g = base + j;
      tile[j] = g < n ? pts[g] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (!active) continue;
    const int lim = min(kKnnTile, n - base);
#pragma unroll 4
    for (int j = 0; j < lim; j++) {
      float4 t = tile[j];
      float dx = __fsub_rn(t.x, qp.x), dy = __fsub_rn(t.y, qp.y), dz = __fsub_rn(t.z, qp.z);
      float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d < worst) {  // targets arrive in increasing index, so ties keep the smaller index: ascending (d2, index)
        int p = k - 1;
        while (p > 0 && ld[(p - 1) * kKnnThreads + tid] > d) {
          ld[p * kKnnThreads + tid] = ld[(p - 1) * kKnnThreads + tid];
          li[p * kKnnThreads + tid] = li[(p - 1) * kKnnThreads + tid];
          p--;
        }
        ld[p * kKnnThreads + tid] = d;
        li[p * kKnnThreads + tid] = base + j;
        worst = ld[(k - 1) * kKnnThreads + tid];
      }
    }
  }
  if (active)
    for (int j = 0; j < k; j++) nbr[(size_t)q * k + j] = li[j * kKnnThreads + tid];
}

// ---------------------------------------------------------------------------------------------------------------
// 3x3 helpers (registers).  Full matrices are row-major m[r*3+c] here.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void inv3_general(const float* m, float* o) {  // Eigen Matrix3f::inverse(): cofactors / det (column 0 expansion)
  float c00 = m[4] * m[8] - m[5] * m[7];
  float c10 = m[7] * m[2] - m[8] * m[1];
  float c20 = m[1] * m[5] - m[2] * m[4];
  float det = (c00 * m[0] + c10 * m[3]) + c20 * m[6];
  float id = 1.0f / det;
  o[0] = c00 * id; o[1] = c10 * id; o[2] = c20 * id;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * id;
  o[4] = (m[8] * m[0] - m[6] * m[2]) * id;
  o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id;
  o[7] = (m[6] * m[1] - m[7] * m[0]) * id;
  o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
__device__ __forceinline__ void mul3(const float* a, const float* b, float* o) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) o[r * 3 + c] = (a[r * 3] * b[c] + a[r * 3 + 1] * b[3 + c]) + a[r * 3 + 2] * b[6 + c];
}
__device__ __forceinline__ float3 cross3(float3 a, float3 b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// Eigen SelfAdjointEigenSolver<Matrix3f>::computeDirect restated (closed-form roots + cross-product eigenvectors);
// s: symmetric, row-major.  evec columns: V[r*3+c].
__device__ __forceinline__ void eig3_extract_kernel(const float* m, float* res /*3*/, float* rep /*3 or null*/) {
  int i0 = 0;
  float best = fabsf(m[0]);
  if (fabsf(m[4]) > best) { best = fabsf(m[4]); i0 = 1; }
  if (fabsf(m[8]) > best) { best = fabsf(m[8]); i0 = 2; }
  int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
  float3 r = make_float3(m[0 * 3 + i0], m[1 * 3 + i0], m[2 * 3 + i0]);
  float3 a = make_float3(m[0 * 3 + i1], m[1 * 3 + i1], m[2 * 3 + i1]);
  float3 b = make_float3(m[0 * 3 + i2], m[1 * 3 + i2], m[2 * 3 + i2]);
  float3 c0 = cross3(r, a), c1 = cross3(r, b);
  float n0 = (c0.x * c0.x + c0.y * c0.y) + c0.z * c0.z;
  float n1 = (c1.x * c1.x + c1.y * c1.y) + c1.z * c1.z;
  if (rep) { rep[0] = r.x; rep[1] = r.y; rep[2] = r.z; }
  if (n0 > n1) {
    float s = sqrtf(n0);
    res[0] = c0.x / s; res[1] = c0.y / s; res[2] = c0.z / s;
  } else {
    float s = sqrtf(n1);
    res[0] = c1.x / s; res[1] = c1.y / s; res[2] = c1.z / s;
  }
}

__device__ void eig3_direct(const float* cov /*row-major sym*/, float* evals, float* V /*V[r*3+c], column c = eigenvector c*/) {
  float s[9];
#pragma unroll
  for (int i = 0; i < 9; i++) s[i] = cov[i];
  // selfadjointView<Lower>
  s[1] = s[3]; s[2] = s[6]; s[5] = s[7];
  float shift = ((s[0] + s[4]) + s[8]) / 3.0f;
  s[0] -= shift; s[4] -= shift; s[8] -= shift;
  float scale = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; i++) scale = fmaxf(scale, fabsf(s[i]));
  if (scale > 0.0f) {
#pragma unroll
    for (int i = 0; i < 9; i++) s[i] = s[i] / scale;
  }
  {  // computeRoots
    const float s_inv3 = 1.0f / 3.0f;
    const float s_sqrt3 = sqrtf(3.0f);
    float c0 = s[0] * s[4] * s[8] + 2.0f * s[3] * s[6] * s[7] - s[0] * s[7] * s[7] - s[4] * s[6] * s[6] - s[8] * s[3] * s[3];
    float c1 = s[0] * s[4] - s[3] * s[3] + s[0] * s[8] - s[6] * s[6] + s[4] * s[8] - s[7] * s[7];
    float c2 = s[0] + s[4] + s[8];
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = fmaxf((c2 * c2_over_3 - c1) * s_inv3, 0.0f);
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = fmaxf(a_over_3 * a_over_3 * a_over_3 - half_b * half_b, 0.0f);
    float rho = sqrtf(a_over_3);
    // atan2 / cos / sin evaluated in double and rounded once: the float results are then the correctly rounded ones
    // whatever libm is underneath (CUDA's atan2f/cosf/sinf are 2-ulp functions; line-like neighbourhoods amplify
    // a 1-ulp difference in theta into a visibly different normal).
    float theta = (float)atan2((double)sqrtf(q), (double)half_b) * s_inv3;
    float cos_theta = (float)cos((double)theta), sin_theta = (float)sin((double)theta);
    evals[0] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    evals[1] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    evals[2] = c2_over_3 + 2.0f * rho * cos_theta;
  }
  const float eps = 1.1920929e-07f;
  float v0[3], v1[3], v2[3];  // eigenvector columns
  if ((evals[2] - evals[0]) <= eps) {
    v0[0] = 1; v0[1] = 0; v0[2] = 0; v1[0] = 0; v1[1] = 1; v1[2] = 0; v2[0] = 0; v2[1] = 0; v2[2] = 1;
  } else {
    float d0 = evals[2] - evals[1];
    float d1 = evals[1] - evals[0];
    bool k_is_2 = d0 > d1;  // k = index of the most distinct eigenvalue, l = the other extreme
    if (k_is_2) d0 = d1;
    float ek = k_is_2 ? evals[2] : evals[0];
    float el = k_is_2 ? evals[0] : evals[2];
    float tmp[9];
#pragma unroll
    for (int i = 0; i < 9; i++) tmp[i] = s[i];
    tmp[0] -= ek; tmp[4] -= ek; tmp[8] -= ek;
    float vk[3], vl[3];
    eig3_extract_kernel(tmp, vk, vl);
    if (d0 <= 2.0f * eps * d1) {
      float dot = (vk[0] * vl[0] + vk[1] * vl[1]) + vk[2] * vl[2];
      float t0 = vl[0] - dot * vl[0], t1 = vl[1] - dot * vl[1], t2 = vl[2] - dot * vl[2];
      float nn = sqrtf((t0 * t0 + t1 * t1) + t2 * t2);
      if (nn > 0.0f) { vl[0] = t0 / nn; vl[1] = t1 / nn; vl[2] = t2 / nn; }
    } else {
#pragma unroll
      for (int i = 0; i < 9; i++) tmp[i] = s[i];
      tmp[0] -= el; tmp[4] -= el; tmp[8] -= el;
      eig3_extract_kernel(tmp, vl, nullptr);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) { v0[i] = k_is_2 ? vl[i] : vk[i]; v2[i] = k_is_2 ? vk[i] : vl[i]; }
    float3 c = cross3(make_float3(v2[0], v2[1], v2[2]), make_float3(v0[0], v0[1], v0[2]));
    float nn = sqrtf((c.x * c.x + c.y * c.y) + c.z * c.z);
    if (nn > 0.0f) { v1[0] = c.x / nn; v1[1] = c.y / nn; v1[2] = c.z / nn; } else { v1[0] = c.x; v1[1] = c.y; v1[2] = c.z; }
  }
#pragma unroll
  for (int r = 0; r < 3; r++) { V[r * 3 + 0] = v0[r]; V[r * 3 + 1] = v1[r]; V[r * 3 + 2] = v2[r]; }
#pragma unroll
  for (int i = 0; i < 3; i++) evals[i] = evals[i] * scale + shift;
}

// covariance_regularization.cu: PLANE :105-116 (V diag(1e-3,1,1) V^-1), MIN_EIG :84-101, FROBENIUS :74-82.
// c: row-major 3x3 in/out.
__device__ void regularize_cov(float* c, int method) {
  if (method == 3 /*PLANE*/ || method == 1 /*MIN_EIG*/) {
    float ev[3], V[9], Vi[9], Vd[9];
    eig3_direct(c, ev, V);
    float l0 = 1e-3f, l1 = 1.0f, l2 = 1.0f;
    if (method == 1) { l0 = fmaxf(1e-3f, ev[0]); l1 = fmaxf(1e-3f, ev[1]); l2 = fmaxf(1e-3f, ev[2]); }
    inv3_general(V, Vi);
#pragma unroll
    for (int r = 0; r < 3; r++) { Vd[r * 3] = V[r * 3] * l0; Vd[r * 3 + 1] = V[r * 3 + 1] * l1; Vd[r * 3 + 2] = V[r * 3 + 2] * l2; }
    mul3(Vd, Vi, c);
  } else if (method == 4 /*FROBENIUS*/) {
    float C[9], Ci[9];
#pragma unroll
    for (int i = 0; i < 9; i++) C[i] = c[i];
    C[0] += 1e-3f; C[4] += 1e-3f; C[8] += 1e-3f;
    inv3_general(C, Ci);
    float nn = 0.0f;  // Frobenius norm, summed in Eigen's (column-major) coefficient order
#pragma unroll
    for (int cidx = 0; cidx < 3; cidx++)
#pragma unroll
      for (int r = 0; r < 3; r++) nn += Ci[r * 3 + cidx] * Ci[r * 3 + cidx];
    nn = sqrtf(nn);
#pragma unroll
    for (int i = 0; i < 9; i++) Ci[i] = Ci[i] / nn;
    inv3_general(Ci, c);
  }
}

__device__ __forceinline__ void store_cov_sym(const float* c /*row-major 3x3*/, float4* __restrict__ covA, float2* __restrict__ covB, int i) {
  // the regularised matrix is symmetric up to float rounding of V*L*V^-1; the packed store keeps the mean of the
  // two triangles (changes the final pose by ~1e-7 m)
  covA[i] = make_float4(c[0], 0.5f * (c[1] + c[3]), 0.5f * (c[2] + c[6]), c[4]);
  covB[i] = make_float2(0.5f * (c[5] + c[7]), c[8]);
}

// Stage 1b: covariance_estimation.cu:26-34 fused with the regulariser.
__global__ void __launch_bounds__(128) k_covariance_knn(const float4* __restrict__ pts, const int* __restrict__ nbr, int n, int k, int method, float4* __restrict__ covA,
                                                       float2* __restrict__ covB) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float mx = 0.f, my = 0.f, mz = 0.f;
  float cxx = 0.f, cxy = 0.f, cxz = 0.f, cyy = 0.f, cyz = 0.f, czz = 0.f;
  const int* row = nbr + (size_t)i * k;
  for (int j = 0; j < k; j++) {
    float4 p = pts[row[j]];
    mx = __fadd_rn(mx, p.x); my = __fadd_rn(my, p.y); mz = __fadd_rn(mz, p.z);
    cxx = __fmaf_rn(p.x, p.x, cxx); cxy = __fmaf_rn(p.x, p.y, cxy); cxz = __fmaf_rn(p.x, p.z, cxz);
    cyy = __fmaf_rn(p.y, p.y, cyy); cyz = __fmaf_rn(p.y, p.z, cyz); czz = __fmaf_rn(p.z, p.z, czz);
  }
  float kf = (float)k;
  mx = __fdiv_rn(mx, kf); my = __fdiv_rn(my, kf); mz = __fdiv_rn(mz, kf);
  float c[9];
  c[0] = __fmaf_rn(-mx, mx, __fdiv_rn(cxx, kf));
  c[1] = c[3] = __fmaf_rn(-mx, my, __fdiv_rn(cxy, kf));
  c[2] = c[6] = __fmaf_rn(-mx, mz, __fdiv_rn(cxz, kf));
  c[4] = __fmaf_rn(-my, my, __fdiv_rn(cyy, kf));
  c[5] = c[7] = __fmaf_rn(-my, mz, __fdiv_rn(cyz, kf));
  c[8] = __fmaf_rn(-mz, mz, __fdiv_rn(czz, kf));
  regularize_cov(c, method);
  store_cov_sym(c, covA, covB, i);
}

// Stage 1b': covariance_estimation_rbf.cu:59-151.  One query per thread, all points streamed through shared memory in
// blocks of 512 like the reference's per-block async transforms; partial sums per 512-block are folded in block order
// (the reference's strided finalisation, :92-114).  The reference pads the cloud to a multiple of 512 with points at the
// origin (:126-129) which pick up weight whenever the query is within max_dist of the origin -- reproduced.
// All nine entries of sum w p p^T are kept ((w p_r) p_c and (w p_c) p_r round differently and the reference's Matrix3f holds
// both).  The weight is exp evaluated in double and rounded once: the reference calls CUDA's 2-ulp expf, which no CPU checker can
// reproduce bit for bit; the correctly rounded value lies within that function's own error bound and makes the stage testable
// bit-for-bit like the rest of this file.
__global__ void __launch_bounds__(128) k_covariance_rbf(const float4* __restrict__ pts, int n, float exp_factor, float max_dist, int method, float4* __restrict__ covA,
                                                       float2* __restrict__ covB) {
  __shared__ float4 tile[kRbfBlock];
  const int tid = threadIdx.x;
  const int q = blockIdx.x * blockDim.x + tid;
  const bool active = q < n;
  float4 x = active ? pts[q] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float max_dist_sq = max_dist * max_dist;
  float sw = 0.f, m[3] = {0.f, 0.f, 0.f}, c[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // c: row-major
  const int nblocks = (n + kRbfBlock - 1) / kRbfBlock;
  for (int b = 0; b < nblocks; b++) {
    __syncthreads();
    for (int j = tid; j < kRbfBlock; j += blockDim.x) {
      int g = b * kRbfBlock + j;
      tile[j] = g < n ? pts[g] : make_float4(0.f, 0.f, 0.f, 0.f);  // padding at the origin, :126-129
    }
    __syncthreads();
    if (!active) continue;
    float psw = 0.f, pm[3] = {0.f, 0.f, 0.f}, pc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < kRbfBlock; j++) {
      float4 p = tile[j];
      float dx = x.x - p.x, dy = x.y - p.y, dz = x.z - p.z;
      float sq = (dx * dx + dy * dy) + dz * dz;
      if (sq > max_dist_sq) continue;
      float w = (float)exp((double)(-exp_factor * sq));
      psw += w;
      float wx = w * p.x, wy = w * p.y, wz = w * p.z;
      pm[0] += wx; pm[1] += wy; pm[2] += wz;
      pc[0] += wx * p.x; pc[1] += wx * p.y; pc[2] += wx * p.z;
      pc[3] += wy * p.x; pc[4] += wy * p.y; pc[5] += wy * p.z;
      pc[6] += wz * p.x; pc[7] += wz * p.y; pc[8] += wz * p.z;
    }
    sw += psw;
#pragma unroll
    for (int d = 0; d < 3; d++) m[d] += pm[d];
#pragma unroll
    for (int d = 0; d < 9; d++) c[d] += pc[d];
  }
  if (!active) return;
  // NormalDistribution::finalize :47-53:  mean = sum/sw ; cov = (cov - mean*sum^T)/sw
  float mean[3] = {m[0] / sw, m[1] / sw, m[2] / sw};
  float cc[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int col = 0; col < 3; col++) cc[r * 3 + col] = (c[r * 3 + col] - mean[r] * m[col]) / sw;
  regularize_cov(cc, method);
  store_cov_sym(cc, covA, covB, q);
}


// ---------------------------------------------------------------------------------------------------------------
// Stage 1a': exact k-NN on a multi-level hash grid, one warp per query.
//
// Why: lidar clouds are surfaces with a density that falls with range; the 20-NN radius spans 0.2 m .. several metres.
// A ladder of L uniform grids (cell size s_l = s_max / 2^(L-1-l), s_max = extent/4) is built in three launches
// (count / allocate / scatter, every point inserted at every level); a query walks the ladder from the finest level
// and stops at the first level where the 3x3x3 (or 5x5x5) block around its cell provably contains its k nearest
// neighbours: every point within r*s of the query lies inside the (2r+1)^3 block, so kth_d2 <= (r*s)^2 certifies it.
// If even the coarsest level cannot certify (tiny clouds, far outliers) the warp scans the whole cloud.
//
// Selection: the warp keeps the k best (d2, index) pairs sorted across its lanes (rank r in lane r%32, register
// r/32); a batch of 32 candidates is loaded coalesced from the cell-sorted copy, lanes whose candidate beats the
// current worst are inserted one at a time with ballot + shuffle (no shared memory, no divergence between queries).
// Ties are ordered by index, so the result is the unique ascending (d2, index) list -- identical to the CPU checker.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGridMaxLevels = 12;
constexpr unsigned long long kGridEmpty = 0xFFFFFFFFFFFFFFFFULL;

struct GridLevel {
  unsigned long long* keys;  // [T] packed cell coordinate, kGridEmpty when free
  int* cnt;                  // [T] points in the cell
  int* start;                // [T] first index of the cell in `sorted`
  int* fill;                 // [T] scatter cursor
  float4* sorted;            // [n] points grouped by cell, original index in .w
  int* pslot;                // [n] table slot of each point (count pass -> scatter pass)
};

struct GridArgs {
  const float4* pts;
  int n, k, L;
  unsigned tmask;            // table size - 1 (same for every level)
  unsigned* bbox_min;        // [3] ordered-uint min xyz (initialised to 0xFFFFFFFF)
  unsigned* bbox_max;        // [3] ordered-uint max xyz (initialised to 0)
  int* level_cursor;         // [L]
  int* query_cursor;         // work counter of the query kernel
  int* heavy_count;          // number of queries deferred to the block-cooperative kernel
  int* heavy_cursor;         // its work counter
  int2* heavy_queue;         // [n] (position in lv[0].sorted, level to resume at)
  GridLevel lv[kGridMaxLevels];
  int* nbr;
};

__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_grid_bbox(const float4* __restrict__ pts, int n, unsigned* __restrict__ bbox_min, unsigned* __restrict__ bbox_max) {
  float lo[3] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
  float hi[3] = {__int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000)};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      atomicMin(&bbox_min[d], f2ord(lo[d]));
      atomicMax(&bbox_max[d], f2ord(hi[d]));
    }
  }
}

struct GridGeom {
  float minx, miny, minz;
  float s_max;
};
__device__ __forceinline__ GridGeom grid_geom(const unsigned* __restrict__ bmin, const unsigned* __restrict__ bmax) {
  GridGeom g;
  g.minx = ord2f(bmin[0]); g.miny = ord2f(bmin[1]); g.minz = ord2f(bmin[2]);
  float ex = ord2f(bmax[0]) - g.minx, ey = ord2f(bmax[1]) - g.miny, ez = ord2f(bmax[2]) - g.minz;
  float e = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-3f));
  g.s_max = e * 0.25f;
  return g;
}
__device__ __forceinline__ float grid_cell_size(const GridGeom& g, int l, int L) { return ldexpf(g.s_max, l - (L - 1)); }
__device__ __forceinline__ int3 grid_cell(const GridGeom& g, float inv_s, float x, float y, float z) {
  return make_int3((int)floorf((x - g.minx) * inv_s), (int)floorf((y - g.miny) * inv_s), (int)floorf((z - g.minz) * inv_s));
}
__device__ __forceinline__ unsigned long long grid_key(int x, int y, int z) {  // 21 bits per axis, offset so that -1 is representable
  return ((unsigned long long)(unsigned)(x + 1024) << 42) | ((unsigned long long)(unsigned)(y + 1024) << 21) | (unsigned long long)(unsigned)(z + 1024);
}
__device__ __forceinline__ unsigned grid_hash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return (unsigned)k;
}

// count pass: thread per (level, point).  Consecutive points are spatially coherent (scan order), so most lanes of a
// warp fall into the same cell, above all on the coarse levels: lanes with equal keys elect a leader that does the
// table probe and one atomicAdd for the group (same-address atomics serialise in L2).
__global__ void k_grid_count(GridArgs a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int l = blockIdx.y;
  const bool active = i < a.n;
  GridGeom g = grid_geom(a.bbox_min, a.bbox_max);
  float inv_s = 1.0f / grid_cell_size(g, l, a.L);
  float4 p = a.pts[active ? i : 0];
  int3 c = grid_cell(g, inv_s, p.x, p.y, p.z);
  unsigned long long key = active ? grid_key(c.x, c.y, c.z) : kGridEmpty;
  GridLevel lv = a.lv[l];
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  unsigned pos = 0;
  if (active && lane == leader) {
    pos = grid_hash(key) & a.tmask;
    for (;;) {
      unsigned long long cur = lv.keys[pos];
      if (cur == kGridEmpty) {
        unsigned long long old = atomicCAS(&lv.keys[pos], kGridEmpty, key);
        cur = (old == kGridEmpty) ? key : old;
      }
      if (cur == key) break;
      pos = (pos + 1) & a.tmask;
    }
    atomicAdd(&lv.cnt[pos], __popc(peers));
  }
  pos = __shfl_sync(0xffffffffu, pos, leader);
  if (active) lv.pslot[i] = (int)pos;
}

// allocate pass: thread per (level, slot): carve the cell's range out of the level's sorted array (one atomic per warp)
__global__ void k_grid_alloc(GridArgs a) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  int l = blockIdx.y;
  GridLevel lv = a.lv[l];
  const int lane = threadIdx.x & 31;
  int c = (t <= a.tmask) ? lv.cnt[t] : 0;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  int base = 0;
  if (lane == 31) base = atomicAdd(&a.level_cursor[l], total);
  base = __shfl_sync(0xffffffffu, base, 31);
  if (c > 0) lv.start[t] = base + incl - c;
}

// scatter pass: thread per (level, point), one cursor atomic per group of lanes sharing a cell
__global__ void k_grid_scatter(GridArgs a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int l = blockIdx.y;
  const bool active = i < a.n;
  GridLevel lv = a.lv[l];
  const int lane = threadIdx.x & 31;
  int slot = active ? lv.pslot[i] : -1 - lane;  // inactive lanes get unique keys
  const unsigned peers = __match_any_sync(0xffffffffu, slot);
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (active && lane == leader) base = lv.start[slot] + atomicAdd(&lv.fill[slot], __popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (!active) return;
  int rank = __popc(peers & ((1u << lane) - 1u));
  float4 p = a.pts[i];
  p.w = __int_as_float(i);
  lv.sorted[base + rank] = p;
}

// ---- warp-level sorted top-k (k <= 64): rank r lives in lane r & 31, register r >> 5.
// An entry is one 64-bit key: (bits of d2) << 32 | index.  d2 >= 0, so unsigned order of the key == ascending (d2, index).
// WIDE=false (k <= 32) keeps a single register set. ----
typedef unsigned long long tkey;
constexpr tkey kKeyInf = 0x7f8000007fffffffULL;  // (+inf, INT_MAX)
__device__ __forceinline__ tkey make_key(float d, int i) { return ((tkey)__float_as_uint(d) << 32) | (unsigned)i; }

struct WarpTopK {
  tkey e0, e1;
  tkey worst;  // entry of rank k-1, broadcast
};

__device__ __forceinline__ void topk_reset(WarpTopK& t) { t.e0 = t.e1 = t.worst = kKeyInf; }

// compare-exchange step of a bitonic network over the 32 lanes: lane keeps the smaller key when keep_min
__device__ __forceinline__ void cmpx(tkey& e, int stride, bool keep_min) {
  tkey o = __shfl_xor_sync(0xffffffffu, e, stride);
  if ((o < e) == keep_min) e = o;  // keys are distinct except (inf, INT_MAX) padding, where either choice is the same
}

// Merge a batch of 32 candidates into the sorted list (k <= 32): bitonic-sort the batch (15 steps), take the element-wise
// minimum with the reversed list (the 32 smallest of the 64, a bitonic sequence), bitonic-merge (5 steps).
__device__ __forceinline__ void topk_merge32(WarpTopK& t, int k, int lane, tkey c) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      bool up = (lane & size) == 0 || size == 32;
      bool lower = (lane & stride) == 0;
      cmpx(c, stride, lower == up);
    }
  }
  if (__all_sync(0xffffffffu, t.e0 == kKeyInf)) {  // (the vote must not sit behind a per-lane short-circuit)
    t.e0 = c;  // empty list: the sorted batch is the list
  } else {
    tkey r = __shfl_sync(0xffffffffu, c, 31 - lane);
    if (r < t.e0) t.e0 = r;
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) cmpx(t.e0, stride, (lane & stride) == 0);
  }
  t.worst = __shfl_sync(0xffffffffu, t.e0, k - 1);
}

// all 32 lanes call this with their candidate (valid=false for padding lanes)
// `bound`: an upper bound on the k-th smallest key known from a previous (smaller) block -- candidates above it cannot be among
// the k nearest and are dropped before they cost a merge
template <bool WIDE>
__device__ __forceinline__ void topk_offer(WarpTopK& t, int k, int lane, bool valid, float cd, int ci, tkey bound = kKeyInf) {
  const tkey c = make_key(cd, ci);
  const bool pass = valid && c < t.worst && c <= bound;
  unsigned m = __ballot_sync(0xffffffffu, pass);
  if (!WIDE && __popc(m) > 3) {  // many entrants: one sort-merge instead of one insertion each
    topk_merge32(t, k, lane, pass ? c : kKeyInf);
    return;
  }
  while (m) {
    int src = __ffs(m) - 1;
    m &= m - 1;
    tkey n = __shfl_sync(0xffffffffu, c, src);
    if (!(n < t.worst)) continue;  // worst moved since the ballot (warp-uniform branch)
    int p = __popc(__ballot_sync(0xffffffffu, t.e0 < n));  // rank of the new entry
    tkey up0 = __shfl_up_sync(0xffffffffu, t.e0, 1);
    const int kr = k - 1;
    if (WIDE) {
      p += __popc(__ballot_sync(0xffffffffu, t.e1 < n));
      tkey up1 = __shfl_up_sync(0xffffffffu, t.e1, 1);
      tkey carry = __shfl_sync(0xffffffffu, t.e0, 31);  // rank 31 -> 32 crosses registers
      int r1 = 32 + lane;
      if (r1 > p) t.e1 = (lane == 0) ? carry : up1;
      else if (r1 == p) t.e1 = n;
    }
    if (lane > p) t.e0 = up0;
    else if (lane == p) t.e0 = n;
    if (WIDE) {
      tkey w0 = __shfl_sync(0xffffffffu, t.e0, kr & 31), w1 = __shfl_sync(0xffffffffu, t.e1, kr & 31);
      t.worst = (kr < 32) ? w0 : w1;
    } else {
      t.worst = __shfl_sync(0xffffffffu, t.e0, kr);
    }
  }
}

__device__ __forceinline__ float topk_worst_d2(const WarpTopK& t) { return __uint_as_float((unsigned)(t.worst >> 32)); }
__device__ __forceinline__ int key_index(tkey e) { return (int)(unsigned)(e & 0xffffffffULL); }

__device__ __forceinline__ float knn_d2(float4 q, float4 t) {  // (dx*dx + dy*dy) + dz*dz, no contraction
  float dx = __fsub_rn(t.x, q.x), dy = __fsub_rn(t.y, q.y), dz = __fsub_rn(t.z, q.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// scan one contiguous run of the cell-sorted array (whole-cloud fallback), four loads in flight
template <bool WIDE>
__device__ __forceinline__ void scan_run(WarpTopK& t, int k, int lane, float4 q, const float4* __restrict__ sorted, int start, int count, tkey bound = kKeyInf) {
  constexpr int U = 4;
  for (int off = 0; off < count; off += 32 * U) {
    float4 c[U];
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      int j = off + u * 32 + lane;
      valid[u] = j < count;
      c[u] = valid[u] ? __ldg(&sorted[start + j]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (off + u * 32 < count) topk_offer<WIDE>(t, k, lane, valid[u], knn_d2(q, c[u]), __float_as_int(c[u].w), bound);
  }
}

__device__ __forceinline__ void probe_cell(const GridLevel& lv, unsigned tmask, int x, int y, int z, int& start, int& count) {
  count = 0;
  start = 0;
  if (x < 0 || y < 0 || z < 0) return;  // no point lies below the bounding-box minimum
  unsigned long long key = grid_key(x, y, z);
  unsigned pos = grid_hash(key) & tmask;
  for (;;) {
    unsigned long long cur = __ldg(&lv.keys[pos]);
    if (cur == key) { start = __ldg(&lv.start[pos]); count = __ldg(&lv.cnt[pos]); return; }
    if (cur == kGridEmpty) return;
    pos = (pos + 1) & tmask;
  }
}

// scan the points of the (up to 32) cells probed by the lanes.  Cells are small (~5 points on the level that wins), so the
// candidates of all cells are flattened into one index space (prefix sum of the counts across lanes) and consumed in
// full 32-wide batches; each lane finds the cell of its candidate with a 5-step binary search over the lanes' prefix
// values.  Four batches (loads) are kept in flight: a sparse query can own thousands of candidates and a single warp
// is latency-bound.
template <bool WIDE>
__device__ __forceinline__ void scan_lane_cells(WarpTopK& t, int k, int lane, float4 q, const float4* __restrict__ sorted, int my_start, int my_count, tkey bound = kKeyInf) {
  int incl = my_count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  const int excl = incl - my_count;
  constexpr int U = 4;
  for (int base = 0; base < total; base += 32 * U) {
    float4 c[U];
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int j = base + u * 32 + lane;
      valid[u] = j < total;
      c[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (base + u * 32 < total) {  // warp-uniform
        int cell = 0;  // largest lane index whose exclusive prefix is <= j (runs of equal prefixes end at the non-empty cell)
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          int e = __shfl_sync(0xffffffffu, excl, (cell + step) & 31);
          if (e <= j) cell += step;
        }
        int st = __shfl_sync(0xffffffffu, my_start, cell);
        int ex = __shfl_sync(0xffffffffu, excl, cell);
        if (valid[u]) c[u] = __ldg(&sorted[st + (j - ex)]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (base + u * 32 < total) topk_offer<WIDE>(t, k, lane, valid[u], knn_d2(q, c[u]), __float_as_int(c[u].w), bound);
    }
  }
}

// the same scans with the batches dealt round-robin to `nw` cooperating warps (warp `wi` takes batches wi, wi+nw, ...)
template <bool WIDE>
__device__ __forceinline__ void scan_lane_cells_strided(WarpTopK& t, int k, int lane, float4 q, const float4* __restrict__ sorted, int my_start, int my_count, int wi, int nw,
                                                        tkey bound = kKeyInf) {
  int incl = my_count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  const int excl = incl - my_count;
  constexpr int U = 4;
  for (int base = wi * 32; base < total; base += 32 * U * nw) {
    float4 c[U];
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int b0 = base + u * 32 * nw;
      const int j = b0 + lane;
      valid[u] = j < total;
      c[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 < total) {
        int cell = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          int e = __shfl_sync(0xffffffffu, excl, (cell + step) & 31);
          if (e <= j) cell += step;
        }
        int st = __shfl_sync(0xffffffffu, my_start, cell);
        int ex = __shfl_sync(0xffffffffu, excl, cell);
        if (valid[u]) c[u] = __ldg(&sorted[st + (j - ex)]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (base + u * 32 * nw < total) topk_offer<WIDE>(t, k, lane, valid[u], knn_d2(q, c[u]), __float_as_int(c[u].w), bound);
  }
}

template <bool WIDE>
__device__ __forceinline__ void scan_run_strided(WarpTopK& t, int k, int lane, float4 q, const float4* __restrict__ sorted, int start, int count, int wi, int nw, tkey bound = kKeyInf) {
  constexpr int U = 4;
  for (int off = wi * 32; off < count; off += 32 * U * nw) {
    float4 c[U];
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      int j = off + u * 32 * nw + lane;
      valid[u] = j < count;
      c[u] = valid[u] ? __ldg(&sorted[start + j]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (off + u * 32 * nw < count) topk_offer<WIDE>(t, k, lane, valid[u], knn_d2(q, c[u]), __float_as_int(c[u].w), bound);
  }
}

// Extension step.  The 3x3x3 block of a level held >= k points, but its k-th distance B exceeds the certificate radius s.  Every
// true neighbour lies within B of q, and while B <= 2 s that ball is inside the 5x5x5 block around q's cell: merging the shell
// cells (5^3 minus the 3^3 already merged) that intersect the ball completes the answer exactly.  On LiDAR surfaces that is a
// handful of points, against the ~4x larger block of the next coarser level a restart would scan.
template <bool WIDE>
__device__ __forceinline__ bool extend_shell(WarpTopK& t, int k, int lane, float4 q, const GridGeom& g, float s, const GridLevel& lv, unsigned tmask, int3 c) {
  const float B = sqrtf(topk_worst_d2(t));
  if (!(B <= 2.f * s * 0.999f)) return false;  // (also rejects an unfilled list)
  const float reach = B + 1e-3f * s;           // slack for the rounding of the cell boundaries
  const float reach2 = reach * reach;
  const float fx = q.x - g.minx, fy = q.y - g.miny, fz = q.z - g.minz;
  for (int r = 0; r < 4; r++) {
    const int idx = r * 32 + lane;
    int st = 0, cn = 0;
    if (idx < 125) {
      const int dx = idx / 25 - 2, dy = (idx / 5) % 5 - 2, dz = idx % 5 - 2;
      if (max(abs(dx), max(abs(dy), abs(dz))) == 2) {
        const int cx = c.x + dx, cy = c.y + dy, cz = c.z + dz;
        const float lx = cx * s, ly = cy * s, lz = cz * s;
        const float ex = fmaxf(fmaxf(lx - fx, fx - (lx + s)), 0.f), ey = fmaxf(fmaxf(ly - fy, fy - (ly + s)), 0.f), ez = fmaxf(fmaxf(lz - fz, fz - (lz + s)), 0.f);
        if (ex * ex + ey * ey + ez * ez <= reach2) probe_cell(lv, tmask, cx, cy, cz, st, cn);
      }
    }
    if (__any_sync(0xffffffffu, cn > 0)) scan_lane_cells<WIDE>(t, k, lane, q, lv.sorted, st, cn);
  }
  return true;
}

constexpr int kKnnGridWarps = 8;      // warps (queries) per block
constexpr int kHeavyCandidates = 768; // blocks with more candidates than this go to the block-cooperative kernel
// nearest-first order of the 3x3x3 block (index = 9*(dx+1) + 3*(dy+1) + (dz+1)): centre, 6 faces, 12 edges, 8 corners
__constant__ unsigned char kBlockOrder[27] = {13, 4, 10, 12, 14, 16, 22, 1, 3, 5, 7, 9, 11, 15, 17, 19, 21, 23, 25, 0, 2, 6, 8, 18, 20, 24, 26};

// Persistent warps pull queries from a global counter (the cost of a query varies by >10x between dense and sparse
// regions, and queries are visited in cell order, so static blocks would leave a long tail).
template <bool WIDE>
__global__ void __launch_bounds__(kKnnGridWarps * 32, 4) k_knn_grid(GridArgs a, int force_bruteforce) {
  const int lane = threadIdx.x & 31;
  const int k = a.k;
  GridGeom g = grid_geom(a.bbox_min, a.bbox_max);
  // nearest-first cell order inside the 3x3x3 block: centre, faces, edges, corners
  int dx = 0, dy = 0, dz = 0;
  if (lane < 27) {
    int o = kBlockOrder[lane];
    dx = o / 9 - 1; dy = (o / 3) % 3 - 1; dz = o % 3 - 1;
  }
  for (;;) {
    int w = 0;
    if (lane == 0) w = atomicAdd(a.query_cursor, 1);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= a.n) return;
    // queries in the finest level's cell order: neighbouring warps touch the same cells
    float4 q = __ldg(&a.lv[0].sorted[w]);
    const int qi = __float_as_int(q.w);
    WarpTopK t;
    topk_reset(t);
    bool done = false, deferred = false;
    tkey bound = kKeyInf;  // k-th key of the last block that failed to certify: an upper bound on the true k-th key
    if (!force_bruteforce) {
      for (int l = 0; l < a.L && !done; l++) {
        const float s = grid_cell_size(g, l, a.L);
        if (bound != kKeyInf) {  // a level whose certificate radius is below the known lower bound s_prev cannot help; one that covers
          const float need = sqrtf(__uint_as_float((unsigned)(bound >> 32)));  // the upper bound certifies for sure: skip in between
          if (l + 1 < a.L && s * 0.999f < need && grid_cell_size(g, l + 1, a.L) * 0.999f <= need) continue;
        }
        const float inv_s = 1.0f / s;
        const GridLevel lv = a.lv[l];
        int3 c = grid_cell(g, inv_s, q.x, q.y, q.z);
        int st = 0, cn = 0;
        if (lane < 27) probe_cell(lv, a.tmask, c.x + dx, c.y + dy, c.z + dz, st, cn);
        int total = cn;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        // A surface sampled at density rho puts ~9 rho s^2 points in the block and ~pi rho s^2 within distance s, so the block
        // can only certify k neighbours when it holds >~ 2.9 k points: below 2.5 k go straight to the coarser level.
        if (total < k || (2 * total < 5 * k && l + 1 < a.L)) continue;
        if (total > kHeavyCandidates) {  // a lone warp is latency-bound: hand big blocks to the block-cooperative kernel
          if (lane == 0) a.heavy_queue[atomicAdd(a.heavy_count, 1)] = make_int2(w, l);
          deferred = true;
          break;
        }
        topk_reset(t);
        scan_lane_cells<WIDE>(t, k, lane, q, lv.sorted, st, cn, bound);
        const float r1 = s * 0.999f;
        if (topk_worst_d2(t) <= r1 * r1) { done = true; break; }  // every point within s of q lies inside the 3x3x3 block
        if (t.worst != kKeyInf) {
          if (extend_shell<WIDE>(t, k, lane, q, g, s, lv, a.tmask, c)) { done = true; break; }
          bound = t.worst;
        }
      }
    }
    if (deferred) continue;
    if (!done) {  // no level could certify the answer (tiny cloud, far outlier): whole cloud, block-cooperative
      if (lane == 0) a.heavy_queue[atomicAdd(a.heavy_count, 1)] = make_int2(w, a.L);
      continue;
    }
    int* row = a.nbr + (size_t)qi * k;
    if (lane < k) row[lane] = key_index(t.e0);
    if (WIDE && 32 + lane < k) row[32 + lane] = key_index(t.e1);
  }
}

// Block-cooperative continuation for the deferred queries: the 8 warps of a block split the candidates of one query
// (warp j takes batches j, j+8, ...), each keeps its own sorted top-k, warp 0 merges the eight lists through shared
// memory and applies the same certificate; a query that still fails moves to the next level, and past the coarsest
// level the block scans the whole cloud.
template <bool WIDE>
__global__ void __launch_bounds__(kKnnGridWarps * 32) k_knn_grid_heavy(GridArgs a) {
  __shared__ tkey se[kKnnGridWarps][64];
  __shared__ int s_next;
  __shared__ int s_done;
  __shared__ tkey s_bound;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int k = a.k;
  GridGeom g = grid_geom(a.bbox_min, a.bbox_max);
  int dx = 0, dy = 0, dz = 0;
  if (lane < 27) {
    int o = kBlockOrder[lane];
    dx = o / 9 - 1; dy = (o / 3) % 3 - 1; dz = o % 3 - 1;
  }
  const int n_heavy = *a.heavy_count;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_next = atomicAdd(a.heavy_cursor, 1);
    __syncthreads();
    const int h = s_next;
    if (h >= n_heavy) return;
    const int2 item = a.heavy_queue[h];
    float4 q = __ldg(&a.lv[0].sorted[item.x]);
    const int qi = __float_as_int(q.w);
    WarpTopK t;
    tkey bound = kKeyInf;
    for (int l = item.y; l <= a.L; l++) {
      topk_reset(t);
      float s = 0.f;
      GridLevel lv = a.lv[0];
      int3 c = make_int3(0, 0, 0);
      if (l < a.L) {
        s = grid_cell_size(g, l, a.L);
        if (bound != kKeyInf) {
          const float need = sqrtf(__uint_as_float((unsigned)(bound >> 32)));
          if (l + 1 < a.L && s * 0.999f < need && grid_cell_size(g, l + 1, a.L) * 0.999f <= need) continue;  // identical in all warps
        }
        const float inv_s = 1.0f / s;
        lv = a.lv[l];
        c = grid_cell(g, inv_s, q.x, q.y, q.z);
        int st = 0, cn = 0;
        if (lane < 27) probe_cell(lv, a.tmask, c.x + dx, c.y + dy, c.z + dz, st, cn);
        int total = cn;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        if (total < k) continue;  // warp-uniform and identical in all warps
        scan_lane_cells_strided<WIDE>(t, k, lane, q, lv.sorted, st, cn, wid, kKnnGridWarps, bound);
      } else {
        scan_run_strided<WIDE>(t, k, lane, q, a.lv[0].sorted, 0, a.n, wid, kKnnGridWarps, bound);
      }
      // merge the per-warp lists in warp 0
      se[wid][lane] = t.e0;
      se[wid][32 + lane] = t.e1;
      __syncthreads();
      if (wid == 0) {
        for (int ww = 1; ww < kKnnGridWarps; ww++) {
          tkey e = se[ww][lane];
          topk_offer<WIDE>(t, k, lane, lane < k, __uint_as_float((unsigned)(e >> 32)), key_index(e));
          if (WIDE) {
            e = se[ww][32 + lane];
            topk_offer<WIDE>(t, k, lane, 32 + lane < k, __uint_as_float((unsigned)(e >> 32)), key_index(e));
          }
        }
        const float r1 = s * 0.999f;
        bool ok = (l == a.L) || (topk_worst_d2(t) <= r1 * r1);
        if (!ok && t.worst != kKeyInf) ok = extend_shell<WIDE>(t, k, lane, q, g, s, lv, a.tmask, c);
        if (ok) {
          int* row = a.nbr + (size_t)qi * k;
          if (lane < k) row[lane] = key_index(t.e0);
          if (WIDE && 32 + lane < k) row[32 + lane] = key_index(t.e1);
        }
        if (lane == 0) { s_done = ok ? 1 : 0; s_bound = t.worst; }
      }
      __syncthreads();
      if (s_done) break;
      if (s_bound != kKeyInf) bound = s_bound;
    }
  }
}

// regulariser over voxel covariances (NDT).  The finalised covariance (S_rc - mean_r * S_c)/n is symmetric only up to rounding;
// Eigen's selfadjointView<Lower> reads the lower triangle, which is what the packed record holds.
__global__ void k_regularize_voxels(VoxelRec* __restrict__ vox, const int* __restrict__ nv_ptr, int method) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= *nv_ptr) return;
  VoxelRec r = vox[v];
  float c[9] = {r.c0.x, r.c0.y, r.c0.z, r.c0.y, r.c0.w, r.c1.x, r.c0.z, r.c1.x, r.c1.y};
  regularize_cov(c, method);
  r.c0 = make_float4(c[0], 0.5f * (c[1] + c[3]), 0.5f * (c[2] + c[6]), c[4]);
  r.c1 = make_float4(0.5f * (c[5] + c[7]), c[8], 0.f, 0.f);
  vox[v] = r;
}

// ---------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------
cudaError_t launch_regularize_voxels(VoxelRec* vox, const int* nv_ptr, int vmax, int method, cudaStream_t stream) {
  k_regularize_voxels<<<(vmax + 127) / 128, 128, 0, stream>>>(vox, nv_ptr, method);
  return cudaGetLastError();
}

size_t knn_smem_bytes(int k) { return sizeof(float4) * kKnnTile + (size_t)k * kKnnThreads * (sizeof(float) + sizeof(int)); }

cudaError_t launch_knn_bruteforce(const float4* pts, int n, int k, int* nbr, cudaStream_t stream) {
  // per device and thread-safe: set on every call (a host-side attribute write; this engine is the A/B legacy path)
  cudaError_t e = cudaFuncSetAttribute(k_knn_bruteforce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_smem_bytes(kMaxK));
  if (e != cudaSuccess) return e;
  k_knn_bruteforce<<<(n + kKnnThreads - 1) / kKnnThreads, kKnnThreads, knn_smem_bytes(k), stream>>>(pts, n, k, nbr);
  return cudaGetLastError();
}

size_t knn_grid_scratch_bytes(int n, int* levels_out, unsigned* table_size_out) {
  int L = 8;
  for (long m = 16384; m * 4 <= (long)n && L < kGridMaxLevels; m *= 4) L++;
  if (n < 4096) L = 6;
  unsigned T = 1024;
  while (T < 2u * (unsigned)n) T <<= 1;
  if (levels_out) *levels_out = L;
  if (table_size_out) *table_size_out = T;
  size_t per_level = (size_t)T * (8 + 4 + 4 + 4) + (size_t)n * (16 + 4);
  return 4096 + (size_t)L * per_level + (size_t)n * 8;
}

// scratch layout:  [0xFF-filled : bbox min (16 B) | keys of all levels]
//                  [zero-filled : bbox max (16 B) | level cursors (64 B) | cnt, fill of all levels]
//                  [uninitialised: start of all levels | sorted copies | pslot]
cudaError_t launch_knn_grid(const float4* pts, int n, int k, int* nbr, unsigned char* scratch, size_t scratch_bytes, int force_bruteforce, int blocks_per_sm_hint, int* launches,
                            cudaStream_t stream) {
  int L;
  unsigned T;
  size_t need = knn_grid_scratch_bytes(n, &L, &T);
  if (scratch_bytes < need || (reinterpret_cast<uintptr_t>(scratch) & 15)) return cudaErrorInvalidValue;
  GridArgs a;
  a.pts = pts; a.n = n; a.k = k; a.L = L; a.tmask = T - 1; a.nbr = nbr;
  unsigned char* p = scratch;
  unsigned char* ff_begin = p;
  a.bbox_min = reinterpret_cast<unsigned*>(p); p += 16;
  for (int l = 0; l < L; l++) { a.lv[l].keys = reinterpret_cast<unsigned long long*>(p); p += (size_t)T * 8; }
  const size_t ff_bytes = (size_t)(p - ff_begin);
  unsigned char* z_begin = p;
  a.bbox_max = reinterpret_cast<unsigned*>(p); p += 16;
  a.level_cursor = reinterpret_cast<int*>(p); p += 64;
  a.query_cursor = a.level_cursor + 15;
  a.heavy_count = a.level_cursor + 14;
  a.heavy_cursor = a.level_cursor + 13;
  for (int l = 0; l < L; l++) {
    a.lv[l].cnt = reinterpret_cast<int*>(p); p += (size_t)T * 4;
    a.lv[l].fill = reinterpret_cast<int*>(p); p += (size_t)T * 4;
  }
  const size_t z_bytes = (size_t)(p - z_begin);
  for (int l = 0; l < L; l++) { a.lv[l].start = reinterpret_cast<int*>(p); p += (size_t)T * 4; }
  for (int l = 0; l < L; l++) { a.lv[l].sorted = reinterpret_cast<float4*>(p); p += (size_t)n * 16; }
  for (int l = 0; l < L; l++) { a.lv[l].pslot = reinterpret_cast<int*>(p); p += (size_t)n * 4; }
  p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
  a.heavy_queue = reinterpret_cast<int2*>(p); p += (size_t)n * 8;
  cudaError_t e;
  if ((e = cudaMemsetAsync(ff_begin, 0xFF, ff_bytes, stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(z_begin, 0, z_bytes, stream)) != cudaSuccess) return e;
  const int nb = (n + 255) / 256;
  k_grid_bbox<<<nb < 592 ? nb : 592, 256, 0, stream>>>(pts, n, a.bbox_min, a.bbox_max);
  k_grid_count<<<dim3(nb, L), 256, 0, stream>>>(a);
  k_grid_alloc<<<dim3((T + 255) / 256, L), 256, 0, stream>>>(a);
  k_grid_scatter<<<dim3(nb, L), 256, 0, stream>>>(a);
  int qblocks = (n + kKnnGridWarps - 1) / kKnnGridWarps;
  // persistent warps pull queries from a counter; 4 blocks (32 warps) per SM by default, 2 when many handles share the GPU
  // (leaves room for other streams' kernels on every SM: +3..6 % aggregate throughput, slower alone)
  const int blocks_per_sm = (blocks_per_sm_hint >= 1 && blocks_per_sm_hint <= 8) ? blocks_per_sm_hint : 4;
  if (qblocks > 148 * blocks_per_sm) qblocks = 148 * blocks_per_sm;  // persistent: queries pulled from a counter
  int hblocks = qblocks < 148 * 4 ? qblocks : 148 * 4;
  if (k <= 32) {
    k_knn_grid<false><<<qblocks, kKnnGridWarps * 32, 0, stream>>>(a, force_bruteforce);
    k_knn_grid_heavy<false><<<hblocks, kKnnGridWarps * 32, 0, stream>>>(a);
  } else {
    k_knn_grid<true><<<qblocks, kKnnGridWarps * 32, 0, stream>>>(a, force_bruteforce);
    k_knn_grid_heavy<true><<<hblocks, kKnnGridWarps * 32, 0, stream>>>(a);
  }
  if (launches) *launches = 6;
  return cudaGetLastError();
}

cudaError_t launch_covariance_knn(const float4* pts, const int* nbr, int n, int k, int method, float4* covA, float2* covB, cudaStream_t stream) {
  k_covariance_knn<<<(n + 127) / 128, 128, 0, stream>>>(pts, nbr, n, k, method, covA, covB);
  return cudaGetLastError();
}

cudaError_t launch_covariance_rbf(const float4* pts, int n, float exp_factor, float max_dist, int method, float4* covA, float2* covB, cudaStream_t stream) {
  k_covariance_rbf<<<(n + 127) / 128, 128, 0, stream>>>(pts, n, exp_factor, max_dist, method, covA, covB);
  return cudaGetLastError();
}

}  // namespace vgicp
