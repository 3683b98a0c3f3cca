// vgicp_prep.cu -- device-side input preparation (include/vgicp_prep_b200.h): near-origin filter + pcl::ApproximateVoxelGrid,
// bit-identical to the serial filter (points and order).  Reference call sites: src/align.cpp:128-147, src/kitti.cpp:80-82,
// src/python/main.cpp:46-62,81-91.  Parity: tests/test_input_prep.py.
//
// Why one block: the filter's 512 history entries are 512 independent sequential chains; thread h of the walk block owns entry h
// and sees the points that hash to it in input order.  Per tile of 2048 staged points every point announces itself to its entry
// (shared-memory atomic append, then the entry sorts its <= 16 tile-local indices: arrival order of the atomics is arbitrary,
// input order is what the filter defines); an entry with more matches in one tile falls back to scanning the staged hashes.
// A flushed centroid is stored at the index of the point that flushed it; a scan over those flags (one block per 1024 points,
// decoupled look-back) gives it the rank the serial filter would have emitted it at, and the entries still holding a voxel at the
// end follow in entry order.
#include <cuda_runtime.h>
#include <stdint.h>

#include <new>
#include <string>

#include "vgicp_prep_b200.h"

namespace {

constexpr int kHist = 512;   // pcl::ApproximateVoxelGrid histsize_ (default; the reference never changes it)
constexpr int kTile = 2048;  // points staged per round of the walk
constexpr int kCap = 16;     // tile-local matches per entry kept in the fast list
constexpr unsigned short kNoEntry = 0xFFFF;

// x y z -> float4 copy + history entry (kNoEntry for a point the near-origin filter drops)
__global__ void k_prep_keys(const float* __restrict__ xyz, size_t stride_f, int n, float inv_leaf, int filter, float4* __restrict__ pts4, unsigned short* __restrict__ entry) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = xyz + (size_t)i * stride_f;
  const float x = p[0], y = p[1], z = p[2];
  // align.cpp:131: squaredNorm() < 1e-3, evaluated as (x*x + y*y) + z*z without contraction (Eigen's unrolled redux order)
  const float sq = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  const bool keep = !(filter && sq < 1e-3f);
  const int ix = (int)floorf(__fmul_rn(x, inv_leaf)), iy = (int)floorf(__fmul_rn(y, inv_leaf)), iz = (int)floorf(__fmul_rn(z, inv_leaf));
  const long long hl = (long long)ix * 7171 + (long long)iy * 3079 + (long long)iz * 4231;
  pts4[i] = make_float4(x, y, z, 0.f);
  entry[i] = keep ? (unsigned short)(hl & (long long)(kHist - 1)) : kNoEntry;
}

struct Entry {  // one history entry, in the registers of its thread
  int ix, iy, iz, count;
  float sx, sy, sz;
};

__device__ __forceinline__ void entry_step(Entry& e, int i, const float4* __restrict__ pts4, float inv_leaf, float4* __restrict__ flushed) {
  const float4 p = pts4[i];
  const int ix = (int)floorf(__fmul_rn(p.x, inv_leaf)), iy = (int)floorf(__fmul_rn(p.y, inv_leaf)), iz = (int)floorf(__fmul_rn(p.z, inv_leaf));
  if (e.count && (e.ix != ix || e.iy != iy || e.iz != iz)) {  // another voxel lands on this entry: emit the centroid held so far
    const float c = (float)e.count;
    flushed[i] = make_float4(__fdiv_rn(e.sx, c), __fdiv_rn(e.sy, c), __fdiv_rn(e.sz, c), 1.f);
    e.count = 0;
    e.sx = e.sy = e.sz = 0.f;
  }
  e.ix = ix; e.iy = iy; e.iz = iz;
  e.count++;
  e.sx = __fadd_rn(e.sx, p.x); e.sy = __fadd_rn(e.sy, p.y); e.sz = __fadd_rn(e.sz, p.z);
}

__global__ void __launch_bounds__(kHist) k_prep_walk(const float4* __restrict__ pts4, const unsigned short* __restrict__ entry, int n, float inv_leaf, float4* __restrict__ flushed,
                                                     float4* __restrict__ tail) {
  __shared__ unsigned short sh_entry[kTile];
  __shared__ int cnt[kHist];
  __shared__ unsigned short list[kHist * kCap];
  const int h = threadIdx.x;
  Entry e;
  e.ix = e.iy = e.iz = e.count = 0;
  e.sx = e.sy = e.sz = 0.f;
  for (int base = 0; base < n; base += kTile) {
    cnt[h] = 0;
    __syncthreads();
    for (int j = h; j < kTile; j += kHist) {
      const int i = base + j;
      const unsigned short t = i < n ? entry[i] : kNoEntry;
      sh_entry[j] = t;
      if (t != kNoEntry) {
        const int p = atomicAdd(&cnt[t], 1);
        if (p < kCap) list[(int)t * kCap + p] = (unsigned short)j;
      }
    }
    __syncthreads();
    const int c = cnt[h];
    if (c > 0 && c <= kCap) {
      unsigned short* mine = &list[h * kCap];  // this thread's segment: sort the tile-local indices ascending (= input order)
      for (int a = 1; a < c; a++) {
        const unsigned short v = mine[a];
        int b = a - 1;
        while (b >= 0 && mine[b] > v) { mine[b + 1] = mine[b]; b--; }
        mine[b + 1] = v;
      }
      for (int q = 0; q < c; q++) entry_step(e, base + (int)mine[q], pts4, inv_leaf, flushed);
    } else if (c > kCap) {  // a run of points on one entry (e.g. the invalid returns at the origin): walk the staged tile
      for (int j = 0; j < kTile; j++)
        if (sh_entry[j] == (unsigned short)h) entry_step(e, base + j, pts4, inv_leaf, flushed);
    }
    __syncthreads();  // the tile's lists are consumed before the next tile overwrites them
  }
  if (e.count) {
    const float c = (float)e.count;
    tail[h] = make_float4(__fdiv_rn(e.sx, c), __fdiv_rn(e.sy, c), __fdiv_rn(e.sz, c), 1.f);
  } else {
    tail[h] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// rank of every flushed centroid = number of flushes at smaller point indices; one block per 1024 points, decoupled look-back over
// epoch-tagged block totals (a block only waits for blocks with a smaller index, which were scheduled before it)
__global__ void __launch_bounds__(1024) k_prep_scatter(const float4* __restrict__ flushed, int n, float* __restrict__ out, size_t cap, unsigned long long* chunk_state,
                                                       unsigned epoch, int* __restrict__ total) {
  __shared__ int warp_sums[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i = blockIdx.x * 1024 + tid;
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n) f = flushed[i];
  const int flag = f.w != 0.f ? 1 : 0;
  int v = flag;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) warp_sums[wid] = v;
  __syncthreads();
  if (wid == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_sums[lane] = w;
    const int block_total = __shfl_sync(0xffffffffu, w, 31);
    volatile unsigned long long* state = chunk_state;
    if (lane == 0) {
      state[blockIdx.x] = ((unsigned long long)epoch << 32) | (unsigned)block_total;
      __threadfence();
    }
    int base = 0;
    for (int p = (int)blockIdx.x - 1 - lane; p >= 0; p -= 32) {
      unsigned long long sv;
      while ((unsigned)((sv = state[p]) >> 32) != epoch) {}
      base += (int)(unsigned)(sv & 0xffffffffULL);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
    if (lane == 0) {
      s_base = base;
      if (blockIdx.x == gridDim.x - 1) *total = base + block_total;
    }
  }
  __syncthreads();
  if (flag) {
    const size_t pos = (size_t)(s_base + (wid > 0 ? warp_sums[wid - 1] : 0) + v - 1);
    if (pos < cap) { out[3 * pos] = f.x; out[3 * pos + 1] = f.y; out[3 * pos + 2] = f.z; }
  }
}

// the entries still holding a voxel, in entry order, behind the flushed centroids
__global__ void __launch_bounds__(kHist) k_prep_tail(const float4* __restrict__ tail, const int* __restrict__ total, float* __restrict__ out, size_t cap, int* __restrict__ n_out) {
  __shared__ int warp_sums[kHist / 32];
  const int h = threadIdx.x, lane = h & 31, wid = h >> 5;
  const float4 f = tail[h];
  const int flag = f.w != 0.f ? 1 : 0;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) warp_sums[wid] = __popc(m);
  __syncthreads();
  int before = 0, all = 0;
  for (int w = 0; w < kHist / 32; w++) {
    const int s = warp_sums[w];
    if (w < wid) before += s;
    all += s;
  }
  const int first = *total;
  if (flag) {
    const size_t pos = (size_t)(first + before + __popc(m & ((1u << lane) - 1u)));
    if (pos < cap) { out[3 * pos] = f.x; out[3 * pos + 1] = f.y; out[3 * pos + 2] = f.z; }
  }
  if (h == 0) *n_out = first + all;
}

template <typename T>
struct DevBuf {  // grow-only device array
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = n + n / 4 + 16;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct vgicp_prep_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  DevBuf<unsigned char> staging;  // host input copied here
  DevBuf<float4> pts4;
  DevBuf<unsigned short> entry;
  DevBuf<float4> flushed;
  DevBuf<float> out;
  DevBuf<unsigned long long> chunk_state;
  unsigned chunk_epoch = 0;
  float4* tail = nullptr;  // [kHist]
  int* d_counts = nullptr;  // [0] flushed total, [1] n_out
  int* h_counts = nullptr;  // pinned
};

namespace {
int fail(vgicp_prep_handle h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
#define PREP_TRY(h, expr)                                                                                   \
  do {                                                                                                      \
    cudaError_t e__ = (expr);                                                                               \
    if (e__ != cudaSuccess) return fail(h, 3, std::string(#expr) + ": " + cudaGetErrorString(e__));         \
  } while (0)
}  // namespace

extern "C" {

int vgicp_prep_create(int device, vgicp_prep_handle* out) {
  if (!out) return 1;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return 5;
  if (device < 0 || device >= ndev) return 1;
  vgicp_prep_context* h = new (std::nothrow) vgicp_prep_context();
  if (!h) return 3;
  h->device = device;
  bool ok = cudaSetDevice(device) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&h->tail, kHist * sizeof(float4)) == cudaSuccess;
  ok = ok && cudaMalloc(&h->d_counts, 2 * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMallocHost(&h->h_counts, 2 * sizeof(int)) == cudaSuccess;
  if (!ok) {
    vgicp_prep_destroy(h);
    return 3;
  }
  *out = h;
  return 0;
}

void vgicp_prep_destroy(vgicp_prep_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  h->staging.release(); h->pts4.release(); h->entry.release(); h->flushed.release(); h->out.release(); h->chunk_state.release();
  if (h->tail) cudaFree(h->tail);
  if (h->d_counts) cudaFree(h->d_counts);
  if (h->h_counts) cudaFreeHost(h->h_counts);
  delete h;
}

const char* vgicp_prep_last_error(vgicp_prep_handle h) { return h ? h->err.c_str() : "null handle"; }

int vgicp_prep_approximate_voxel_grid(vgicp_prep_handle h, const float* xyz, size_t n, size_t stride_bytes, int on_device, float leaf, int remove_near_origin, float* out_xyz,
                                      size_t cap, int out_on_device, size_t* n_out) {
  if (!h) return 1;
  if (!n_out) return fail(h, 1, "approximate_voxel_grid: n_out is null");
  *n_out = 0;
  if (n == 0) return 0;
  if (!xyz || !out_xyz) return fail(h, 1, "approximate_voxel_grid: null buffer");
  if (stride_bytes < 12 || stride_bytes % 4) return fail(h, 1, "approximate_voxel_grid: stride must be >= 12 and a multiple of 4");
  if (!(leaf > 0.f)) return fail(h, 1, "approximate_voxel_grid: leaf size must be positive");
  if (n > (size_t)0x7fffffff - 4096) return fail(h, 1, "approximate_voxel_grid: too many points");
  PREP_TRY(h, cudaSetDevice(h->device));
  const int ni = (int)n;
  const float* d_in = xyz;
  if (!on_device) {
    PREP_TRY(h, h->staging.reserve(n * stride_bytes));
    PREP_TRY(h, cudaMemcpyAsync(h->staging.p, xyz, (n - 1) * stride_bytes + 12, cudaMemcpyHostToDevice, h->stream));
    d_in = reinterpret_cast<const float*>(h->staging.p);
  }
  PREP_TRY(h, h->pts4.reserve(n));
  PREP_TRY(h, h->entry.reserve(n));
  PREP_TRY(h, h->flushed.reserve(n));
  float* d_out = out_xyz;
  const size_t out_cap = out_on_device ? cap : n;  // the filter never emits more points than it reads
  if (!out_on_device) {
    PREP_TRY(h, h->out.reserve(3 * n));
    d_out = h->out.p;
  }
  const int chunks = (ni + 1023) / 1024;
  {
    const unsigned long long* before = h->chunk_state.p;
    PREP_TRY(h, h->chunk_state.reserve((size_t)chunks));
    if (h->chunk_state.p != before) {  // fresh allocation: no stale epoch may match
      PREP_TRY(h, cudaMemsetAsync(h->chunk_state.p, 0, h->chunk_state.cap * sizeof(unsigned long long), h->stream));
      h->chunk_epoch = 0;
    }
    if (++h->chunk_epoch == 0) {
      PREP_TRY(h, cudaMemsetAsync(h->chunk_state.p, 0, h->chunk_state.cap * sizeof(unsigned long long), h->stream));
      h->chunk_epoch = 1;
    }
  }
  const float inv_leaf = 1.0f / leaf;
  k_prep_keys<<<(ni + 255) / 256, 256, 0, h->stream>>>(d_in, stride_bytes / 4, ni, inv_leaf, remove_near_origin ? 1 : 0, h->pts4.p, h->entry.p);
  PREP_TRY(h, cudaMemsetAsync(h->flushed.p, 0, n * sizeof(float4), h->stream));
  k_prep_walk<<<1, kHist, 0, h->stream>>>(h->pts4.p, h->entry.p, ni, inv_leaf, h->flushed.p, h->tail);
  k_prep_scatter<<<chunks, 1024, 0, h->stream>>>(h->flushed.p, ni, d_out, out_cap, h->chunk_state.p, h->chunk_epoch, h->d_counts);
  k_prep_tail<<<1, kHist, 0, h->stream>>>(h->tail, h->d_counts, d_out, out_cap, h->d_counts + 1);
  PREP_TRY(h, cudaGetLastError());
  PREP_TRY(h, cudaMemcpyAsync(h->h_counts, h->d_counts, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  PREP_TRY(h, cudaStreamSynchronize(h->stream));
  const size_t m = (size_t)h->h_counts[1];
  *n_out = m;
  if (m > cap) return fail(h, 1, "approximate_voxel_grid: output capacity too small (n points always suffice)");
  if (!out_on_device && m > 0) {
    PREP_TRY(h, cudaMemcpyAsync(out_xyz, d_out, m * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    PREP_TRY(h, cudaStreamSynchronize(h->stream));
  }
  return 0;
}

}  // extern "C"
