// vgicp_batch.cpp -- include/vgicp_batch_b200.h: a pool of vgicp handles on one device and worker threads that pull pairs from a
// shared counter.  Host-only C++ over the public C ABI (vgicp_b200.h); every pair is one vgicp_register call on one handle, i.e.
// the body of the reference's benchmark loop (src/align.cpp:72-81), so results equal the sequential loop's whatever the interleaving.
#include "vgicp_batch_b200.h"

#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

struct vgicp_batch_context {
  int device = 0;
  std::vector<vgicp_handle> handles;
  std::string err;
  std::mutex err_mutex;
};

namespace {
enum { kOk = 0, kInvalidArgument = 1, kCuda = 3 };  // vgicp_b200.h status values

int fail(vgicp_batch_handle b, int code, const std::string& msg) {
  if (b) {
    std::lock_guard<std::mutex> lock(b->err_mutex);
    b->err = msg;
  }
  return code;
}
}  // namespace

extern "C" {

int vgicp_batch_create(int device, int n_streams, vgicp_batch_handle* out) {
  if (!out) return kInvalidArgument;
  *out = nullptr;
  if (n_streams < 1 || n_streams > 64) return kInvalidArgument;
  vgicp_batch_context* b = new (std::nothrow) vgicp_batch_context();
  if (!b) return kCuda;
  b->device = device;
  for (int s = 0; s < n_streams; s++) {
    vgicp_handle h = nullptr;
    const int rc = vgicp_create(device, &h);
    if (rc != kOk) {
      vgicp_batch_destroy(b);
      return rc;
    }
    b->handles.push_back(h);
  }
  *out = b;
  return kOk;
}

void vgicp_batch_destroy(vgicp_batch_handle b) {
  if (!b) return;
  for (vgicp_handle h : b->handles) vgicp_destroy(h);
  delete b;
}

const char* vgicp_batch_last_error(vgicp_batch_handle b) { return b ? b->err.c_str() : "null handle"; }

int vgicp_batch_num_streams(vgicp_batch_handle b) { return b ? (int)b->handles.size() : 0; }

int vgicp_batch_configure(vgicp_batch_handle b, double resolution, int neighbor_search_method, double radius) {
  if (!b) return kInvalidArgument;
  for (vgicp_handle h : b->handles) {
    int rc = vgicp_set_resolution(h, resolution);
    if (rc == kOk) rc = vgicp_set_neighbor_search_method(h, neighbor_search_method, radius);
    if (rc != kOk) return fail(b, rc, vgicp_last_error(h));
  }
  return kOk;
}

int vgicp_batch_register(vgicp_batch_handle b, size_t n_pairs, const float* const* target_xyz, const size_t* n_target, const float* const* source_xyz, const size_t* n_source,
                         size_t stride_bytes, int on_device, int k, int regularization_method, const double* guesses, const vgicp_lsq_params* params,
                         vgicp_align_result* results, float* const* aligned_out) {
  if (!b) return kInvalidArgument;
  if (n_pairs == 0) return kOk;
  if (!target_xyz || !n_target || !source_xyz || !n_source || !results) return fail(b, kInvalidArgument, "batch_register: null argument");
  std::atomic<size_t> next{0};
  std::atomic<int> first_error{kOk};
  auto worker = [&](vgicp_handle h) {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= n_pairs || first_error.load() != kOk) return;
      int rc = vgicp_register(h, target_xyz[i], n_target[i], source_xyz[i], n_source[i], stride_bytes, on_device, k, regularization_method, guesses ? guesses + 16 * i : nullptr,
                              params, &results[i]);
      if (rc == kOk && aligned_out && aligned_out[i]) rc = vgicp_transform_source(h, results[i].T, aligned_out[i], n_source[i], 12);
      if (rc != kOk) {
        int expected = kOk;
        if (first_error.compare_exchange_strong(expected, rc)) fail(b, rc, std::string("pair ") + std::to_string(i) + ": " + vgicp_last_error(h));
        return;
      }
    }
  };
  const size_t n_workers = b->handles.size() < n_pairs ? b->handles.size() : n_pairs;
  std::vector<std::thread> threads;
  threads.reserve(n_workers);
  try {
    for (size_t s = 1; s < n_workers; s++) threads.emplace_back(worker, b->handles[s]);
  } catch (...) {  // could not start a thread: the ones running (and this one) still drain the counter
  }
  worker(b->handles[0]);
  for (std::thread& t : threads) t.join();
  return first_error.load();
}

}  // extern "C"
