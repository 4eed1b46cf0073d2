// clock_shim.h — force-included (-include) into the build of the reference's viewpoint_manager.cpp.
// uncNeighVps seeds eight std::default_random_engine objects from
// std::chrono::system_clock::now().time_since_epoch().count() (viewpoint_manager.cpp:962-977), which
// makes the reference non-deterministic.  This header pulls in every standard header that mentions
// system_clock first and then renames the identifier for the rest of the translation unit to a clock
// whose n-th now() returns `seed + n`: exactly the injected-seed scheme of the oracle and of the
// CUDA library (engine k of call c is seeded with seed + 8c + k).  TEST INFRASTRUCTURE.
#pragma once
#include <chrono>
#include <condition_variable>
#include <future>
#include <iostream>
#include <memory>
#include <mutex>
#include <random>
#include <thread>
#include <unordered_map>

namespace std {
namespace chrono {
struct fcvm_shim_clock {
  typedef std::chrono::nanoseconds duration;
  typedef duration::rep rep;
  typedef duration::period period;
  typedef std::chrono::time_point<fcvm_shim_clock, duration> time_point;
  static unsigned long long& base() { static unsigned long long b = 0; return b; }
  static unsigned long long& calls() { static unsigned long long c = 0; return c; }
  static time_point now() { return time_point(duration((rep)(base() + calls()++))); }
};
}  // namespace chrono
}  // namespace std
#define system_clock fcvm_shim_clock
