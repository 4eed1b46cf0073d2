// fcvm_hostutil.h — small host-side utilities shared by the translation units of libfcvm.so:
// error reporting behind fcvm_last_error(), CUDA call checking, device / pinned buffers, a thread pool loop.
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
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
// Host staging buffer for D2H / H2D copies.  Page-locking costs ~0.4 ms per MB on the B200 boxes (600 MB: 240 ms, measured,
// profiles/pin_cost.py) against 23 ms saved per 600 MB copied, so a large buffer starts pageable and is page-locked in place
// (cudaHostRegister) only once it has been used a few times; small buffers are page-locked at once.
template <class T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  bool pinned = false, registered = false;
  int uses = 0;
  static constexpr size_t kPinAtOnceBytes = 8u << 20;
  static constexpr int kPinAfterUses = 3;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    free();
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    if (bytes <= kPinAtOnceBytes) {
      cudaError_t e = cudaMallocHost(&p, bytes);
      if (e != cudaSuccess) { p = nullptr; return e; }
      pinned = true;
    } else {
      void* q = nullptr;
      const size_t padded = (bytes + 4095) / 4096 * 4096;
      if (posix_memalign(&q, 4096, padded) != 0) return cudaErrorMemoryAllocation;
      p = (T*)q;
      // First touch in parallel: a fresh 600 MB region filled by one staged D2H copy takes ~300 ms of page faults
      // (1.6 GB/s, measured), eight threads pre-fault it in ~25 ms.
      const int nt = 8;
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t)
        th.emplace_back([=] {
          const size_t lo = padded / nt * t / 4096 * 4096, hi = t == nt - 1 ? padded : padded / nt * (t + 1) / 4096 * 4096;
          for (size_t o = lo; o < hi; o += 4096) ((volatile char*)q)[o] = 0;
        });
      for (auto& x : th) x.join();
    }
    cap = n;
    return cudaSuccess;
  }
  // call once per use as a copy target / source: page-locks a large buffer that keeps being reused
  void note_use() {
    if (pinned || !p) return;
    if (++uses >= kPinAfterUses) {
      const size_t bytes = (std::max<size_t>(cap, 1) * sizeof(T) + 4095) / 4096 * 4096;
      if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess) pinned = registered = true;
      else (void)cudaGetLastError();      // stays pageable: slower copies, still correct
    }
  }
  void free() {
    if (p) {
      if (registered) { cudaHostUnregister(p); ::free(p); }
      else if (pinned) cudaFreeHost(p);
      else ::free(p);
    }
    p = nullptr; cap = 0; pinned = registered = false; uses = 0;
  }
};

}  // namespace fcvm
