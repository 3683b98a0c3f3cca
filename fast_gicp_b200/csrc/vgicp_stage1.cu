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
__global__ void __launch_bounds__(128) k_covariance_rbf(const float4* __restrict__ pts, int n, float exp_factor, float max_dist, int method, float4* __restrict__ covA,
                                                       float2* __restrict__ covB) {
  __shared__ float4 tile[kRbfBlock];
  const int tid = threadIdx.x;
  const int q = blockIdx.x * blockDim.x + tid;
  const bool active = q < n;
  float4 x = active ? pts[q] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float max_dist_sq = max_dist * max_dist;
  float sw = 0.f, m[3] = {0.f, 0.f, 0.f}, c[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int nblocks = (n + kRbfBlock - 1) / kRbfBlock;
  for (int b = 0; b < nblocks; b++) {
    __syncthreads();
    for (int j = tid; j < kRbfBlock; j += blockDim.x) {
      int g = b * kRbfBlock + j;
      tile[j] = g < n ? pts[g] : make_float4(0.f, 0.f, 0.f, 0.f);  // padding at the origin, :126-129
    }
    __syncthreads();
    if (!active) continue;
    float psw = 0.f, pm[3] = {0.f, 0.f, 0.f}, pc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < kRbfBlock; j++) {
      float4 p = tile[j];
      float dx = x.x - p.x, dy = x.y - p.y, dz = x.z - p.z;
      float sq = (dx * dx + dy * dy) + dz * dz;
      if (sq > max_dist_sq) continue;
      float w = expf(-exp_factor * sq);
      psw += w;
      float wx = w * p.x, wy = w * p.y, wz = w * p.z;
      pm[0] += wx; pm[1] += wy; pm[2] += wz;
      pc[0] += wx * p.x; pc[1] += wx * p.y; pc[2] += wx * p.z; pc[3] += wy * p.y; pc[4] += wy * p.z; pc[5] += wz * p.z;
    }
    sw += psw;
#pragma unroll
    for (int d = 0; d < 3; d++) m[d] += pm[d];
#pragma unroll
    for (int d = 0; d < 6; d++) c[d] += pc[d];
  }
  if (!active) return;
  // NormalDistribution::finalize :47-53:  mean = sum/sw ; cov = (cov - mean*sum^T)/sw
  float mean[3] = {m[0] / sw, m[1] / sw, m[2] / sw};
  float cc[9];
  cc[0] = (c[0] - mean[0] * m[0]) / sw; cc[1] = (c[1] - mean[0] * m[1]) / sw; cc[2] = (c[2] - mean[0] * m[2]) / sw;
  cc[3] = (c[1] - mean[1] * m[0]) / sw; cc[4] = (c[3] - mean[1] * m[1]) / sw; cc[5] = (c[4] - mean[1] * m[2]) / sw;
  cc[6] = (c[2] - mean[2] * m[0]) / sw; cc[7] = (c[4] - mean[2] * m[1]) / sw; cc[8] = (c[5] - mean[2] * m[2]) / sw;
  regularize_cov(cc, method);
  store_cov_sym(cc, covA, covB, q);
}


// ---------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------
size_t knn_smem_bytes(int k) { return sizeof(float4) * kKnnTile + (size_t)k * kKnnThreads * (sizeof(float) + sizeof(int)); }

cudaError_t launch_knn_bruteforce(const float4* pts, int n, int k, int* nbr, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_knn_bruteforce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_smem_bytes(kMaxK));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_knn_bruteforce<<<(n + kKnnThreads - 1) / kKnnThreads, kKnnThreads, knn_smem_bytes(k), stream>>>(pts, n, k, nbr);
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
