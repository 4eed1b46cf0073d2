// fcvm_hostutil.h — small host-side utilities shared by the translation units of libfcvm.so:
// error reporting behind fcvm_last_error(), CUDA call checking, device / pinned buffers, a thread pool loop.
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
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

// Environment switches (they select profiling output and A/B variants; none changes results).  env() is read once per
// process; a handle takes its own copy of the A/B switches when it is created (EnvSettings::now()), so one process can
// hold handles of both variants - tests/test_gpu_pose_sums.py compares them.
struct EnvSettings {
  int profile = 0;          // FCVM_PROFILE: 1 = wall-clock split after fcvm_update, 2 = one line per evaluator batch, 3 = + marks
  int pipe_subs = 0;        // FCVM_PIPE_SUBS: sub-batches of the streamed evaluator paths (0 = default per path)
  bool host_sums = false;   // FCVM_HOST_SUMS=1: pitch / yaw sums of getPose entirely on the host (the round-1 path; A/B and cross-check)
  bool own_sweep = false;   // FCVM_OWN_SWEEP=1: ownership by the ordered row sweep instead of the per-point arg-max (single GPU only)
  bool no_pfree = false;    // FCVM_NO_PFREE=1: k_eval marches the viewpoint-side half-ray even where the summed-area table certifies it free
  bool plain_alloc = false; // FCVM_PLAIN_ALLOC=1: handle buffers from cudaMalloc instead of the stream-ordered pool (A/B)
  std::vector<int> pipe_fracs;   // FCVM_PIPE_FRACS=40,25,15,12,8: sub-batch sizes of the CSR path in per cent (tuning; the default is built in)
  EnvSettings() {
    if (const char* e = std::getenv("FCVM_PIPE_FRACS")) {
      int v = 0; bool any = false;
      for (const char* c = e;; ++c) {
        if (*c >= '0' && *c <= '9') { v = v * 10 + (*c - '0'); any = true; }
        else { if (any && v > 0) pipe_fracs.push_back(v); v = 0; any = false; if (!*c) break; }
      }
    }
    if (const char* e = std::getenv("FCVM_PROFILE")) profile = atoi(e);
    if (const char* e = std::getenv("FCVM_PIPE_SUBS")) pipe_subs = std::max(1, atoi(e));
    if (const char* e = std::getenv("FCVM_HOST_SUMS")) host_sums = atoi(e) != 0;
    if (const char* e = std::getenv("FCVM_OWN_SWEEP")) own_sweep = atoi(e) != 0;
    if (const char* e = std::getenv("FCVM_NO_PFREE")) no_pfree = atoi(e) != 0;
    if (const char* e = std::getenv("FCVM_PLAIN_ALLOC")) plain_alloc = atoi(e) != 0;
  }
};
inline EnvSettings env_now() { return EnvSettings(); }
inline const EnvSettings& env() {
  static const EnvSettings e;
  return e;
}

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

// A small persistent worker pool: run(n, f) calls f(i) for i in [0, n) on the pool's threads and the caller's, in blocks
// handed out by an atomic counter.  parallel_for above spawns its threads per call (~0.3 ms for 16), which is too much for
// the ~1 ms pieces of the pose-sum pipeline.
class WorkerPool {
 public:
  explicit WorkerPool(int threads) {
    for (int t = 1; t < threads; ++t) workers_.emplace_back([this] { loop(); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  template <class F>
  void run(int64_t n, int64_t block, F&& f) {
    if (n <= 0) return;
    if (workers_.empty() || n <= block) {
      for (int64_t i = 0; i < n; ++i) f(i);
      return;
    }
    std::function<void(int64_t)> fn = std::ref(f);
    {
      std::lock_guard<std::mutex> g(m_);
      fn_ = &fn; n_ = n; block_ = block; next_.store(0); active_ = (int)workers_.size(); ++epoch_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this] { return active_ == 0; });
    fn_ = nullptr;
  }

 private:
  void work() {
    for (;;) {
      const int64_t i0 = next_.fetch_add(block_);
      if (i0 >= n_) break;
      const int64_t e = std::min(n_, i0 + block_);
      for (int64_t i = i0; i < e; ++i) (*fn_)(i);
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return stop_ || epoch_ != seen; });
        if (stop_) return;
        seen = epoch_;
      }
      work();
      {
        std::lock_guard<std::mutex> g(m_);
        if (--active_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int64_t)>* fn_ = nullptr;
  int64_t n_ = 0, block_ = 1;
  std::atomic<int64_t> next_{0};
  int active_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

struct Stopwatch {
  double& acc;
  std::chrono::steady_clock::time_point t0;
  explicit Stopwatch(double& a) : acc(a), t0(std::chrono::steady_clock::now()) {}
  ~Stopwatch() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// Device buffer that only grows.  Owns its allocation (freed by the destructor on the device that was current when it
// was made: every public entry point selects the handle's device first); movable, not copyable.
// the library's allocation stream of the current device (see DevBuf::reserve); nullptr if it cannot be created
inline cudaStream_t alloc_stream() {
  static std::mutex m;
  static cudaStream_t streams[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { (void)cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> g(m);
  if (!streams[dev] && cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); streams[dev] = nullptr; }
  return streams[dev];
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { free(); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; }
    return *this;
  }
  ~DevBuf() { free(); }
  // From the device's default stream-ordered pool (its release threshold is raised in ensure_device, so it keeps what is
  // freed): a first-time buffer is carved out of memory the pool already owns instead of costing a driver allocation
  // (~0.1-0.3 ms each, ~35 of them on a first updateViewpoints).  The allocation is made on a private stream of the library
  // (one per device, never used for anything else) and that stream is synchronised, so the buffer is usable on every stream
  // at once and nothing of the host application is waited for; cudaFree of such a pointer synchronises the device like before.
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaErrorNotSupported;
    cudaStream_t as = env().plain_alloc ? nullptr : alloc_stream();
    if (as) {
      e = cudaMallocAsync((void**)&p, std::max<size_t>(n, 1) * sizeof(T), as);
      if (e == cudaSuccess) e = cudaStreamSynchronize(as);
    }
    if (e != cudaSuccess) {                   // e.g. a device without memory pools: the plain allocator
      (void)cudaGetLastError();
      p = nullptr;
      e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
    }
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void free() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
// Device buffer from the stream-ordered allocator: growing it neither waits for, nor stalls, work on other streams
// (cudaFree is device-synchronising, cudaFreeAsync is not), which the pose-sum pipeline relies on while the next
// evaluator sub-batch is running.  Grows with 25 % slack.  The device's default pool keeps what is freed (its release
// threshold is raised in ensure_device).
template <class T>
struct AsyncBuf {
  T* p = nullptr;
  size_t cap = 0;
  AsyncBuf() = default;
  AsyncBuf(const AsyncBuf&) = delete;
  AsyncBuf& operator=(const AsyncBuf&) = delete;
  ~AsyncBuf() { if (p) cudaFree(p); }
  cudaError_t reserve(size_t n, cudaStream_t s) {
    if (n <= cap) return cudaSuccess;
    if (p) { cudaError_t e = cudaFreeAsync(p, s); if (e != cudaSuccess) return e; }
    p = nullptr; cap = 0;
    const size_t want = n + n / 4 + 1024;
    cudaError_t e = cudaMallocAsync((void**)&p, want * sizeof(T), s);
    if (e == cudaSuccess) cap = want;
    return e;
  }
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
  PinBuf() = default;
  PinBuf(const PinBuf&) = delete;
  PinBuf& operator=(const PinBuf&) = delete;
  PinBuf(PinBuf&& o) noexcept : p(o.p), cap(o.cap), pinned(o.pinned), registered(o.registered), uses(o.uses) { o.p = nullptr; o.cap = 0; o.pinned = o.registered = false; o.uses = 0; }
  PinBuf& operator=(PinBuf&& o) noexcept {
    if (this != &o) {
      free();
      p = o.p; cap = o.cap; pinned = o.pinned; registered = o.registered; uses = o.uses;
      o.p = nullptr; o.cap = 0; o.pinned = o.registered = false; o.uses = 0;
    }
    return *this;
  }
  ~PinBuf() { free(); }
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
