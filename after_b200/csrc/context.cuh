// Handle-wide plumbing: host copies of loaded tensors, tracked device allocations, launch counter,
// stream bridge (work always runs on a library-owned stream so that it can be graph-captured even
// when the caller passes the legacy default stream).
#pragma once
#include <atomic>
#include <map>
#include <string>
#include <vector>
#include "common.cuh"
#include "../../include/after_b200.h"

namespace after {

extern std::atomic<int64_t> g_launches;
#define AFTER_COUNT_LAUNCH() (::after::g_launches.fetch_add(1, std::memory_order_relaxed))

// Per-kernel-class device timing for the roofline report: CUDA events recorded on the launching (work) stream
// around every launch of the profiled classes.  While enabled, graph replay is bypassed (events inside a
// captured graph cannot be timed), so the numbers are per-launch durations of the very same kernels.
enum KernelClass {
  KC_TAP_GEMM_TC = 0, KC_TAP_GEMM_SIMT = 1, KC_ATTENTION = 2, KC_ROW_NORM = 3, KC_ACT_OPERAND = 4, KC_PQMF = 5,
  KC_OTHER = 6, KC_MLP_FUSED = 7, KC_COUNT = 8
};

struct Profiler {
  bool on = false;
  struct Rec { int cls; cudaEvent_t a, b; double flops, bytes; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    AFTER_CUDA_CHECK(cudaEventCreate(&e));
    return e;
  }
  void reset() {
    for (auto& r : recs) { pool.push_back(r.a); pool.push_back(r.b); }
    recs.clear();
  }
  // sums over finished records of one class; caller synchronises the device first
  void read(int cls, int64_t* n, double* ms, double* flops, double* bytes) {
    *n = 0; *ms = 0; *flops = 0; *bytes = 0;
    for (auto& r : recs) {
      if (r.cls != cls) continue;
      float t = 0.f;
      AFTER_CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
      *n += 1; *ms += t; *flops += r.flops; *bytes += r.bytes;
    }
  }
};
extern Profiler g_prof;

struct ProfScope {
  cudaStream_t st;
  bool active;
  ProfScope(int cls, cudaStream_t s, double flops, double bytes) : st(s), active(g_prof.on) {
    if (!active) return;
    Profiler::Rec r{cls, g_prof.get(), g_prof.get(), flops, bytes};
    cudaEventRecord(r.a, st);
    g_prof.recs.push_back(r);
  }
  ~ProfScope() {
    if (active) cudaEventRecord(g_prof.recs.back().b, st);
  }
};

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

using TensorMap = std::map<std::string, HostTensor>;

struct Arena {
  std::vector<void*> ptrs;
  size_t bytes = 0;
  template <typename T>
  T* alloc(size_t n) {
    void* p = nullptr;
    size_t b = ((n * sizeof(T) + 255) / 256) * 256;
    if (b == 0) b = 256;
    AFTER_CUDA_CHECK(cudaMalloc(&p, b));
    ptrs.push_back(p);
    bytes += b;
    return reinterpret_cast<T*>(p);
  }
  template <typename T>
  T* upload(const std::vector<T>& v) {
    T* p = alloc<T>(v.size());
    AFTER_CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
    bytes = 0;
  }
};

inline const HostTensor& need(const TensorMap& m, const std::string& key, std::initializer_list<int64_t> shape) {
  auto it = m.find(key);
  if (it == m.end()) throw Error(AFTER_EMISSING, "missing tensor '" + key + "'");
  std::vector<int64_t> want(shape);
  if (it->second.shape != want) {
    std::string got;
    for (auto s : it->second.shape) got += std::to_string(s) + ",";
    std::string exp;
    for (auto s : want) exp += std::to_string(s) + ",";
    throw Error(AFTER_ESHAPE, "tensor '" + key + "' has shape (" + got + ") expected (" + exp + ")");
  }
  return it->second;
}

// weight = g * v / ||v||  with the norm over all dims but 0 (torch weight_norm dim=0); SimpleNetsStream.py:84-92.
// Computed in double on the host at load time.
inline std::vector<float> fold_weight_norm(const HostTensor& v, const HostTensor& g) {
  const int64_t d0 = v.shape[0];
  const int64_t inner = v.numel() / d0;
  std::vector<float> w(v.data.size());
  for (int64_t i = 0; i < d0; ++i) {
    double s = 0;
    for (int64_t j = 0; j < inner; ++j) { double x = v.data[i * inner + j]; s += x * x; }
    // torch computes the norm in fp32; match its rounding of the scale factor
    float nrm = (float)std::sqrt(s);
    float sc = g.data[i] / nrm;
    for (int64_t j = 0; j < inner; ++j) w[i * inner + j] = v.data[i * inner + j] * sc;
  }
  return w;
}

struct StreamBridge {
  cudaStream_t work = nullptr;
  cudaEvent_t e_in = nullptr, e_out = nullptr;
  void init() {
    AFTER_CUDA_CHECK(cudaStreamCreateWithFlags(&work, cudaStreamNonBlocking));
    AFTER_CUDA_CHECK(cudaEventCreateWithFlags(&e_in, cudaEventDisableTiming));
    AFTER_CUDA_CHECK(cudaEventCreateWithFlags(&e_out, cudaEventDisableTiming));
  }
  void enter(cudaStream_t user) {
    AFTER_CUDA_CHECK(cudaEventRecord(e_in, user));
    AFTER_CUDA_CHECK(cudaStreamWaitEvent(work, e_in, 0));
  }
  void exit(cudaStream_t user) {
    AFTER_CUDA_CHECK(cudaEventRecord(e_out, work));
    AFTER_CUDA_CHECK(cudaStreamWaitEvent(user, e_out, 0));
  }
  void destroy() {
    if (e_in) cudaEventDestroy(e_in);
    if (e_out) cudaEventDestroy(e_out);
    if (work) cudaStreamDestroy(work);
    e_in = e_out = nullptr;
    work = nullptr;
  }
};

}  // namespace after
