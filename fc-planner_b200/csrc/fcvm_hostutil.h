// fcvm_hostutil.h — small host-side utilities shared by the translation units of libfcvm.so:
// error reporting behind fcvm_last_error(), CUDA call checking, device / pinned buffers, a thread pool loop.
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/fcvm.h"

namespace fcvm {

int fail(int code, const std::string& msg);     // records the message for fcvm_last_error() (thread-local), returns code
const char* last_error_cstr();

#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fcvm::fail(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ? FCVM_ENODEVICE : FCVM_ECUDA, \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                             \
  } while (0)

template <class F>
static void parallel_for(int64_t n, int threads, F&& f) {
  if (threads <= 1 || n < 2 * threads) {
    for (int64_t i = 0; i < n; ++i) f(i);
    return;
  }
  std::atomic<int64_t> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t)
    th.emplace_back([&] {
      for (;;) {
        const int64_t i0 = next.fetch_add(64);
        if (i0 >= n) break;
        for (int64_t i = i0; i < std::min(n, i0 + 64); ++i) f(i);
      }
    });
  for (auto& t : th) t.join();
}

struct Stopwatch {
  double& acc;
  std::chrono::steady_clock::time_point t0;
  explicit Stopwatch(double& a) : acc(a), t0(std::chrono::steady_clock::now()) {}
  ~Stopwatch() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void free() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <class T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMallocHost(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void free() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace fcvm
