// peaks.cu — the two denominators SURVEY.md 8(d) leaves "to be measured once": L2 streaming-read bandwidth and the
// non-tensor FP64 rate (DADD / DMUL / DFMA), plus the dependent-load latency that bounds the DDA march (L2 hit, 8-byte
// loads).  Stand-alone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/scripts/peaks profiles/scripts/peaks.cu
// Prints one JSON object.  CUDA events on the launching stream, best of 5 after a warm-up.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); std::exit(1); } } while (0)

__global__ void k_read(const uint4* __restrict__ p, size_t n, int reps, unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const uint4 v = __ldcg(p + i);        // cache at L2 only
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678u) *sink = acc;
}

template <int OP>
__global__ void k_fp64(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { x0 = __dadd_rn(x0, a); x1 = __dadd_rn(x1, a); x2 = __dadd_rn(x2, a); x3 = __dadd_rn(x3, a); x4 = __dadd_rn(x4, a); x5 = __dadd_rn(x5, a); x6 = __dadd_rn(x6, a); x7 = __dadd_rn(x7, a); }
    if (OP == 1) { x0 = __dmul_rn(x0, b); x1 = __dmul_rn(x1, b); x2 = __dmul_rn(x2, b); x3 = __dmul_rn(x3, b); x4 = __dmul_rn(x4, b); x5 = __dmul_rn(x5, b); x6 = __dmul_rn(x6, b); x7 = __dmul_rn(x7, b); }
    if (OP == 2) { x0 = __fma_rn(x0, b, a); x1 = __fma_rn(x1, b, a); x2 = __fma_rn(x2, b, a); x3 = __fma_rn(x3, b, a); x4 = __fma_rn(x4, b, a); x5 = __fma_rn(x5, b, a); x6 = __fma_rn(x6, b, a); x7 = __fma_rn(x7, b, a); }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// pointer chase: next[i] is a random permutation cycle over n 8-byte slots; one thread, dependent __ldcg loads
__global__ void k_chase(const unsigned long long* __restrict__ next, int steps, unsigned long long* out, long long* cycles) {
  unsigned long long i = *out;           // continue where the previous launch stopped: a walk through slots not touched before
  const long long t0 = clock64();
  for (int s = 0; s < steps; ++s) i = __ldcg(next + i);
  const long long t1 = clock64();
  *out = i; *cycles = t1 - t0;
}

template <class F>
static float best_ms(F&& f, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  const int sms = pr.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  unsigned* sink; CK(cudaMalloc(&sink, 4));
  std::printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"sm_clock_khz\": %d", pr.name, sms, pr.l2CacheSize, clk_khz);
  // ---- L2 streaming read: working sets well inside the L2 (one partition / both partitions) and well outside (HBM)
  const size_t sizes_mb[] = {16, 32, 64, 96, 4096};
  std::printf(", \"read_gbs\": {");
  for (int k = 0; k < 5; ++k) {
    const size_t bytes = sizes_mb[k] << 20, n = bytes / 16;
    uint4* p; CK(cudaMalloc(&p, bytes)); CK(cudaMemset(p, 1, bytes));
    const int reps = sizes_mb[k] >= 1024 ? 2 : 64;
    const float ms = best_ms([&] { k_read<<<sms * 8, 512>>>(p, n, reps, sink); });
    std::printf("%s\"%zu_MB\": %.1f", k ? ", " : "", sizes_mb[k], (double)bytes * reps / (ms * 1e-3) / 1e9);
    CK(cudaFree(p));
  }
  std::printf("}");
  // ---- FP64 pipe
  double* out; CK(cudaMalloc(&out, (size_t)sms * 16 * 256 * sizeof(double)));
  const int iters = 1 << 14;
  const double ops = (double)sms * 16 * 256 * 8.0 * iters;
  const float ms_add = best_ms([&] { k_fp64<0><<<sms * 16, 256>>>(out, iters, 1e-9, 1.0000001); });
  const float ms_mul = best_ms([&] { k_fp64<1><<<sms * 16, 256>>>(out, iters, 1e-9, 1.0000001); });
  const float ms_fma = best_ms([&] { k_fp64<2><<<sms * 16, 256>>>(out, iters, 1e-9, 1.0000001); });
  std::printf(", \"fp64\": {\"dadd_tops\": %.2f, \"dmul_tops\": %.2f, \"dfma_tflops\": %.2f, \"dadd_per_clk_per_sm\": %.1f}", ops / (ms_add * 1e-3) / 1e12,
              ops / (ms_mul * 1e-3) / 1e12, 2 * ops / (ms_fma * 1e-3) / 1e12, ops / (ms_add * 1e-3) / ((double)clk_khz * 1e3) / sms);
  // ---- dependent-load latency, 8-byte loads: 1 MB (L1-resident), 32 MB and 100 MB (L2), 2 GB (HBM)
  std::printf(", \"dependent_load_ns\": {");
  const size_t chase_mb[] = {32, 100, 2048};
  for (int k = 0; k < 3; ++k) {
    const size_t n = (chase_mb[k] << 20) / 8;
    std::vector<unsigned long long> h(n);
    // a single cycle through all slots with a large odd stride (co-prime with n = power-of-two-ish sizes are not guaranteed: use a LCG walk)
    std::vector<unsigned> perm(n);
    for (size_t i = 0; i < n; ++i) perm[i] = (unsigned)i;
    unsigned long long s = 88172645463325252ull;
    for (size_t i = n - 1; i > 0; --i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; const size_t j = s % (i + 1); std::swap(perm[i], perm[j]); }
    for (size_t i = 0; i < n; ++i) h[perm[i]] = perm[(i + 1) % n];
    unsigned long long *d, *o; long long* cyc;
    CK(cudaMalloc(&d, n * 8)); CK(cudaMalloc(&o, 8)); CK(cudaMalloc(&cyc, 8));
    CK(cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice));
    const int steps = 200000;
    CK(cudaMemset(o, 0, 8));
    // sets that fit the L2: two full warm walks leave them resident; the 2 GB set: every launch walks fresh slots (HBM)
    const bool fits = chase_mb[k] <= 100;
    if (fits) for (size_t w = 0; w < 2 * n / steps + 2; ++w) { k_chase<<<1, 1>>>(d, steps, o, cyc); }
    CK(cudaDeviceSynchronize());
    k_chase<<<1, 1>>>(d, steps, o, cyc); CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    std::printf("%s\"%zu_MB\": %.1f", k ? ", " : "", chase_mb[k], (double)c / steps / ((double)clk_khz * 1e3) * 1e9);
    CK(cudaFree(d)); CK(cudaFree(o)); CK(cudaFree(cyc));
  }
  std::printf("}}\n");
  return 0;
}
