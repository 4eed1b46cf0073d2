// fcvm_cov.cu — pin-hole coverage evaluation on sm_100a: kernels + C ABI.
//
// Replaces, for whole batches of poses, the three functions of
// FC-Planner/src/hierarchical_coverage_planner/src/hcplanner.cpp that brute-force every
// path / trajectory pose against every point of the full cloud:
//   CameraModelProjection :1196-1211    PinHoleCamera :1128-1194    CoverageEvaluation :1032-1126
// and the path-coverage loop around PinHoleCamera (:250-271).
//
// Reference algorithm per pose: scan the cloud in index order; a point inside the FOV is projected to
// a pixel; the pixel keeps the point with the smallest camera distance (strict '<', so the earliest
// index wins ties); the visible set is the set of pixel occupants.  CoverageEvaluation additionally
// clears cover_state of every occupant that is displaced during the scan (:1085), which makes the
// per-pose `visibleCloud` depend on the scan order.
//
// Re-design (no sort, no per-pose image in shared memory):
//   k_cov_scan   persistent CTAs pull poses from a counter.  A CTA owns a z-buffer slot in HBM
//                (u64 distance bits + occupant id per pixel, 'empty' = all ones) that only ever
//                holds the touched pixels of one pose and is restored through a touched list, so
//                nothing is cleared per pose.  The cloud is walked in 1024-point chunks in index
//                order; chunks whose bounding box is out of range or entirely behind one FOV
//                plane are skipped (exact: every point of such a chunk fails insideFOV).  Inside a
//                chunk all points are tested at once; the sequential "holds the pixel at the time
//                of its scan" predicate of the reference is evaluated in parallel as
//                    held(i) = d_i < zbuf_before_chunk[pix]  and  no j < i in the chunk with the
//                              same pixel and d_j <= d_i
//                (the second clause walks a per-pixel chain kept in a shared-memory hash table; the chain
//                has one member unless two points of the chunk really share a pixel).
//                Held points atomicMin the pixel; the unique held point that equals the new
//                minimum becomes the occupant and marks the previous one displaced; the other held
//                points mark themselves displaced.  Output: one visible bitset row and (optionally)
//                one displaced bitset row per pose.
//   k_cov_sum /  the order-dependent part of CoverageEvaluation across poses, per 32-point word column:
//   k_cov_apply      cs &= ~displaced_k;  group_k = visible_k & ~cs;  cs |= visible_k;  ever |= visible_k
//                run as a two-pass segmented scan over the poses (row steps compose, see the kernels).
//   rows -> CSR  with k_popc_rows / k_exscan / k_compact_rows of the visibility path.
//
// Arithmetic: FP64, -fmad=false, only + - * / sqrt floor on the device, Eigen's 3-term reduction order
// (SURVEY.md A.2); sin / cos / tan of the pose and camera run on the host (fcvm_geom.h).
// No CPU fallback: without a CUDA device every evaluating call returns FCVM_ENODEVICE.
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>

#include "fcvm_geom.h"
#include "fcvm_hostutil.h"
#include "fcvm_kernels.h"

namespace fcvm {

constexpr int kCovThreads = 256;
#ifndef COV_PER
#define COV_PER 1           // points per thread per chunk
#endif
constexpr int kCovPer = COV_PER;
constexpr int kCovChunk = kCovThreads * kCovPer;  // 256 points, walked in index order; one bounding box each
#ifndef COV_CTAS
#define COV_CTAS 3          // resident CTAs per SM (74 KB of shared memory each)
#endif
constexpr int kCovCtasPerSm = COV_CTAS;
#ifndef COV_GROUP
#define COV_GROUP 2         // live chunks tested per barrier
#endif
constexpr int kCovGroup = COV_GROUP;
#ifndef COV_BATCH
#define COV_BATCH 2048      // projected points a CTA collects (over several chunks) before it resolves their pixels
#endif
constexpr int kCovBatch = COV_BATCH;
constexpr int kCovPerBatch = kCovBatch / kCovThreads;
#ifndef COV_HASH
#define COV_HASH (2 * COV_BATCH)
#endif
constexpr int kCovHash = COV_HASH;                // pixel -> chain table slots (open addressing, load <= 0.5); a power of two
static_assert((kCovHash & (kCovHash - 1)) == 0 && kCovHash > kCovBatch && kCovHash <= 32768, "table size");
constexpr int kCovLivePass = 8192;                // chunks culled per pass (bitmap in shared memory)
constexpr unsigned long long kEmpty = ~0ull;
static_assert(kCovBatch % kCovThreads == 0 && kCovBatch >= 2 * kCovGroup * kCovChunk && kCovBatch <= 32767, "batch size");

struct alignas(16) CovPose {
  VpRec rec;          // position + 4 world FOV normals (setPose_PY)
  double M[9];        // R_p.inverse() * R_y.inverse(), row-major (CameraModelProjection :1199-1204)
  double pad[3];
};

struct CovConst {
  double fx, fy;      // hcplanner/fx, fy
  double ldc0, ldc1;  // leftdown = (-xRange, -yRange)
  double xRes, yRes;
  int xPixel, yPixel;
};

struct CovArgs {
  FovConst fov;
  CovConst cam;
  const float4* points;
  long long P;
  const float* chunk_box;         // [nchunks][6] = min xyz, max xyz of each 1024-point chunk
  int nchunks;
  const CovPose* poses;
  int m;
  uint32_t* vis_rows;             // m x words, zeroed by the caller
  uint32_t* disp_rows;            // m x words or null
  long long words;
  unsigned long long* zdist;      // [slots][npix]
  int* zidx;                      // [slots][npix]
  int* touched;                   // [slots][npix]
  long long npix;
  int* next_pose;                 // work counter (zeroed by the caller)
  unsigned long long* counters;   // [0] projections [1] out-of-image [2] chunks scanned [3] held points whose pixel was shared inside the chunk
};

// (int)floor(v) as x86-64 cvttsd2si delivers it: out-of-range and NaN become INT_MIN
__device__ __forceinline__ int floor_to_int(double v) {
  const double f = floor(v);
  if (!(f >= -2147483648.0 && f <= 2147483647.0)) return INT_MIN;
  return (int)f;
}

size_t cov_smem_bytes() {
  return (size_t)kCovBatch * (sizeof(double) + 2 * sizeof(int) + 2 * sizeof(short)) + (size_t)kCovHash * 2 * sizeof(int);
}

// Round 2: the projected points of SEVERAL chunks are collected in shared memory and their pixels resolved together.  The
// predicate "holds the pixel at the time the reference scans it" only compares point indices, so it applies to any batch that
// is a contiguous piece of the scan.  A pose whose projections all fit one batch (the usual case: ~860 per pose on the shipped
// trajectories) is resolved entirely in shared memory - the z-buffer slot in HBM, its four dependent L2 round trips and three
// of the four barriers per chunk (round 1: 23 chunks per pose) are gone; a pose that overflows the batch takes the z-buffer
// path for all its batches.
__global__ void __launch_bounds__(kCovThreads, kCovCtasPerSm) k_cov_scan(const __grid_constant__ CovArgs a) {
  extern __shared__ __align__(16) unsigned char cov_smem[];
  double* s_dist = reinterpret_cast<double*>(cov_smem);                 // camera distance of each collected point
  int* s_pix = reinterpret_cast<int*>(s_dist + kCovBatch);              // its pixel
  int* s_idx = s_pix + kCovBatch;                                       // its index in the cloud (scan position)
  int* t_key = s_idx + kCovBatch;                                       // open-addressing table pixel -> chain head
  int* t_head = t_key + kCovHash;
  short* s_next = reinterpret_cast<short*>(t_head + kCovHash);          // chain of the batch's points that share a pixel
  short* s_bucket = s_next + kCovBatch;                                 // the table slot of the point's pixel
  __shared__ uint32_t s_live[kCovLivePass / 32];                        // surviving-chunk bitmap of a pass
  __shared__ int s_pose, s_ntouched, s_count;
  __shared__ CovPose s_ps;

  const int tid = threadIdx.x, lane = tid & 31;
  unsigned long long* zdist = a.zdist + (long long)blockIdx.x * a.npix;
  int* zidx = a.zidx + (long long)blockIdx.x * a.npix;
  int* touched = a.touched + (long long)blockIdx.x * a.npix;
  const FovConst& fc = a.fov;
  const CovConst& cc = a.cam;
  unsigned int n_proj = 0, n_out = 0, n_chunks = 0, n_slow = 0;

  for (int i = tid; i < kCovHash; i += kCovThreads) { t_key[i] = -1; t_head[i] = -1; }

  for (;;) {
    __syncthreads();
    if (tid == 0) { s_pose = atomicAdd(a.next_pose, 1); s_ntouched = 0; s_count = 0; }
    __syncthreads();
    const int k = s_pose;
    if (k >= a.m) break;

    // the pose record lives in shared memory (broadcast reads): 28 doubles per thread would cost the occupancy
    if (tid < (int)(sizeof(CovPose) / 16))
      reinterpret_cast<double2*>(&s_ps)[tid] = __ldg(reinterpret_cast<const double2*>(a.poses + k) + tid);
    __syncthreads();
    const CovPose& ps = s_ps;
    const VpRec& rec = ps.rec;
    uint32_t* vrow = a.vis_rows + (long long)k * a.words;
    uint32_t* drow = a.disp_rows ? a.disp_rows + (long long)k * a.words : nullptr;
    bool spilled = false;          // the pose overflowed a batch: its pixels live in the z-buffer slot

    // CameraModelProjection (hcplanner.cpp:1196-1211) of the n collected points: pixel (-1 = not in the image), camera distance,
    // and the point joins its pixel's chain
    auto project = [&](int n) {
#pragma unroll 1
      for (int u = 0; u < kCovPerBatch; ++u) {
        const int e = u * kCovThreads + tid;
        if (e >= n) break;
        const float4 q = __ldg(a.points + s_idx[e]);
        const double dx = (double)q.x - rec.px, dy = (double)q.y - rec.py, dz = (double)q.z - rec.pz;
        const double dd = sqrt((dx * dx + dy * dy) + dz * dz);
        const double pc0 = (ps.M[0] * dx + ps.M[1] * dy) + ps.M[2] * dz;
        const double pc1 = (ps.M[3] * dx + ps.M[4] * dy) + ps.M[5] * dz;
        const double pc2 = (ps.M[6] * dx + ps.M[7] * dy) + ps.M[8] * dz;
        const double camX = cc.fx * pc1 / (1000.0 * pc0) - cc.ldc0;
        const double camY = cc.fy * pc2 / (1000.0 * pc0) - cc.ldc1;
        const int xi = floor_to_int(camX / cc.xRes), yi = floor_to_int(camY / cc.yRes);
        n_proj++;
        int px = -1;
        if (xi >= 0 && yi >= 0) {
          if (xi >= cc.xPixel || yi >= cc.yPixel) {
            n_out++;
          } else {
            px = xi * cc.yPixel + yi;
            s_dist[e] = dd;
            int hb = (int)(((unsigned)px * 0x9E3779B1u) >> 16) & (kCovHash - 1);
            for (;;) {
              const int old = atomicCAS(&t_key[hb], -1, px);
              if (old == -1 || old == px) break;
              hb = (hb + 1) & (kCovHash - 1);
            }
            s_next[e] = (short)atomicExch(&t_head[hb], e);
            s_bucket[e] = (short)hb;
          }
        }
        s_pix[e] = px;
      }
      __syncthreads();
    };

    // Resolve the n collected points against the z-buffer slot (the pose has more projections than one batch holds).
    auto flush_zbuffer = [&](int n) {
      project(n);
      unsigned held = 0u, fresh = 0u;          // bit u: entry u holds its pixel / the pixel was empty before this batch
#pragma unroll 1
      for (int u = 0; u < kCovPerBatch; ++u) {
        const int e = u * kCovThreads + tid;
        if (e >= n) break;
        const int px = s_pix[e];
        if (px < 0) continue;
        const double dd = s_dist[e];
        const unsigned long long before = __ldcg(zdist + px);
        if (before == kEmpty) fresh |= 1u << u;
        if ((unsigned long long)__double_as_longlong(dd) < before) {
          bool h = true;
          const int me = s_idx[e];
          int j = t_head[s_bucket[e]];
          n_slow += (j != e || s_next[e] >= 0);
          for (; j >= 0; j = s_next[j])
            if (s_idx[j] < me && s_dist[j] <= dd) { h = false; break; }
          if (h) held |= 1u << u;
        }
      }
      __syncthreads();
      if (tid == 0) s_count = 0;
      // the pixel's new minimum
#pragma unroll 1
      for (int u = 0; u < kCovPerBatch; ++u) {
        const int e = u * kCovThreads + tid;
        if ((held >> u) & 1u) atomicMin(zdist + s_pix[e], (unsigned long long)__double_as_longlong(s_dist[e]));
      }
      __syncthreads();
      // occupant hand-over; displaced marks; the chain table is wiped
#pragma unroll 1
      for (int u = 0; u < kCovPerBatch; ++u) {
        const int e = u * kCovThreads + tid;
        if (e >= n) break;
        const int px = s_pix[e];
        if (px < 0) continue;
        const int hb = s_bucket[e];
        t_key[hb] = -1;
        t_head[hb] = -1;
        if (!((held >> u) & 1u)) continue;
        const int me = s_idx[e];
        if (__ldcg(zdist + px) == (unsigned long long)__double_as_longlong(s_dist[e])) {
          if ((fresh >> u) & 1u) {
            touched[atomicAdd(&s_ntouched, 1)] = px;
          } else if (drow) {
            const int old = __ldcg(zidx + px);
            atomicOr(drow + (old >> 5), 1u << (old & 31));
          }
          __stcg(zidx + px, me);
        } else if (drow) {
          atomicOr(drow + (me >> 5), 1u << (me & 31));
        }
      }
      __syncthreads();      // the table is clean before the next chunk inserts
    };

    // Resolve the n collected points of a pose that fits one batch, in shared memory alone.  Per pixel chain: a point holds
    // the pixel when it is scanned iff no earlier point of the chain is at least as near; of the holders, the one no point of
    // the chain is strictly nearer than is the occupant at the end of the scan (visible), the others were displaced.
    auto flush_local = [&](int n) {
      project(n);
#pragma unroll 1
      for (int u = 0; u < kCovPerBatch; ++u) {
        const int e = u * kCovThreads + tid;
        if (e >= n) break;
        if (s_pix[e] < 0) continue;
        const double dd = s_dist[e];
        const int me = s_idx[e];
        bool held = true, last = true;
        int j = t_head[s_bucket[e]];
        n_slow += (j != e || s_next[e] >= 0);
        for (; j >= 0; j = s_next[j]) {
          const double dj = s_dist[j];
          if (s_idx[j] < me && dj <= dd) { held = false; break; }
          if (dj < dd) last = false;
        }
        if (held) {
          if (last) atomicOr(vrow + (me >> 5), 1u << (me & 31));
          else if (drow) atomicOr(drow + (me >> 5), 1u << (me & 31));
        }
      }
      __syncthreads();
      if (tid == 0) s_count = 0;
#pragma unroll 1
      for (int u = 0; u < kCovPerBatch; ++u) {
        const int e = u * kCovThreads + tid;
        if (e >= n || s_pix[e] < 0) continue;
        const int hb = s_bucket[e];
        t_key[hb] = -1;
        t_head[hb] = -1;
      }
    };

    for (int c0 = 0; c0 < a.nchunks; c0 += kCovLivePass) {
      const int cn = min(kCovLivePass, a.nchunks - c0);
      // ---- chunk cull: out of range, or the whole box behind one FOV plane ------------------
      __syncthreads();
      for (int cb = 0; cb < cn; cb += kCovThreads) {
        const int c = cb + tid;
        bool live = false;
        if (c < cn) {
          const float* b = a.chunk_box + (long long)(c0 + c) * 6;
          const double lx = (double)__ldg(b + 0) - rec.px, ly = (double)__ldg(b + 1) - rec.py, lz = (double)__ldg(b + 2) - rec.pz;
          const double hx = (double)__ldg(b + 3) - rec.px, hy = (double)__ldg(b + 4) - rec.py, hz = (double)__ldg(b + 5) - rec.pz;
          const double cx = fmax(fmax(lx, -hx), 0.0), cy = fmax(fmax(ly, -hy), 0.0), cz = fmax(fmax(lz, -hz), 0.0);
          live = !((cx * cx + cy * cy) + cz * cz > fc.md2_hi * 1.000001);
          // largest d.n over the box; a margin far above rounding keeps the skip exact
          const double reach = fmax(fabs(lx), fabs(hx)) + fmax(fabs(ly), fabs(hy)) + fmax(fabs(lz), fabs(hz));
          const double margin = 1e-9 * (reach + 1.0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double best = fmax(lx * rec.n[q][0], hx * rec.n[q][0]) + fmax(ly * rec.n[q][1], hy * rec.n[q][1]) +
                                fmax(lz * rec.n[q][2], hz * rec.n[q][2]);
            live = live && !(best < -margin);
          }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, live);
        if (lane == 0) s_live[(cb + tid) >> 5] = mask;
      }
      __syncthreads();

      // ---- surviving chunks in index order ---------------------------------------------------------------------------
      const int nw = (cn + 31) / 32;
      int w = 0;
      uint32_t bits = s_live[0];
      auto next_chunk = [&]() -> int {
        while (bits == 0u) {
          if (++w >= nw) return -1;
          bits = s_live[w];
        }
        const int c = c0 + w * 32 + (__ffs(bits) - 1);
        bits &= bits - 1;
        return c;
      };
      for (;;) {
        // kCovGroup live chunks per barrier: culling stays fine-grained, the barrier (and the batch-overflow check) is per group
        int cs[kCovGroup];
        float4 pts[kCovGroup][kCovPer];
#pragma unroll
        for (int g = 0; g < kCovGroup; ++g) {
          cs[g] = next_chunk();
          if (cs[g] >= 0) {
            const long long base = (long long)cs[g] * kCovChunk;
#pragma unroll
            for (int u = 0; u < kCovPer; ++u) {
              const long long gi = base + u * kCovThreads + tid;
              if (gi < a.P) pts[g][u] = __ldg(a.points + gi);
            }
          }
        }
        if (cs[0] < 0) break;
        // FOV test of the chunks' points; the ones inside join the batch by index.  Their projection (three divisions and a
        // square root each) waits for the batch to be resolved, where it runs on all threads evenly: done here, the few
        // threads whose point is inside kept the whole CTA waiting at the barrier below (27 % of the stall samples).
#pragma unroll
        for (int g = 0; g < kCovGroup; ++g) {
          if (cs[g] < 0) continue;
          n_chunks += (tid == 0);
          const long long base = (long long)cs[g] * kCovChunk;
#pragma unroll
          for (int u = 0; u < kCovPer; ++u) {
            const long long gi = base + u * kCovThreads + tid;
            if (gi < a.P) {
              const float4 q = pts[g][u];
              if (inside_fov(rec, fc, (double)q.x, (double)q.y, (double)q.z)) s_idx[atomicAdd(&s_count, 1)] = (int)gi;   // < kCovBatch, see below
            }
          }
        }
        __syncthreads();
        const int n = s_count;
        if (n > kCovBatch - kCovGroup * kCovChunk) {       // the next group might not fit
          spilled = true;
          flush_zbuffer(n);
        }
      }
    }

    // ---- end of the scan ----------------------------------------------------------------------
    __syncthreads();
    const int n = s_count;
    if (!spilled) {
      if (n > 0) flush_local(n);
    } else {
      if (n > 0) flush_zbuffer(n);
      // occupants -> visible row; z-buffer slot back to 'empty'
      __syncthreads();
      const int nt = s_ntouched;
      for (int t = tid; t < nt; t += kCovThreads) {
        const int p = touched[t];
        const int id = __ldcg(zidx + p);
        atomicOr(vrow + (id >> 5), 1u << (id & 31));
        __stcg(zdist + p, kEmpty);
      }
    }
  }

  // counters: one atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_proj += __shfl_xor_sync(0xffffffffu, n_proj, o);
    n_out += __shfl_xor_sync(0xffffffffu, n_out, o);
    n_chunks += __shfl_xor_sync(0xffffffffu, n_chunks, o);
    n_slow += __shfl_xor_sync(0xffffffffu, n_slow, o);
  }
  if (lane == 0 && a.counters) {
    if (n_proj) atomicAdd(a.counters + 0, (unsigned long long)n_proj);
    if (n_out) atomicAdd(a.counters + 1, (unsigned long long)n_out);
    if (n_chunks) atomicAdd(a.counters + 2, (unsigned long long)n_chunks);
    if (n_slow) atomicAdd(a.counters + 3, (unsigned long long)n_slow);
  }
}

// The order-dependent part of CoverageEvaluation across poses, per 32-point word column c:
//     cs &= ~displaced_k;  group_k = visible_k & ~cs;  cs |= visible_k;  ever |= visible_k      (k in pose order)
// One row step is the map cs -> (cs & ~d) | v; such maps compose: (cs & ~A) | B followed by a row gives
// A' = A | d, B' = (B & ~d) | v.  So the poses are cut into segments that run in parallel:
//   k_cov_sum    one thread per (segment, column): the segment's (A, B); its union of visible bits -> ever (atomicOr)
//   k_cov_apply  one thread per (segment, column): cs at the segment start from the summaries of the earlier
//                segments, then the rows of the segment for real (group rows overwrite the displaced rows);
//                the last segment stores cs for the next batch of the call.
__global__ void __launch_bounds__(128) k_cov_sum(const uint32_t* __restrict__ vis, const uint32_t* __restrict__ disp, int m, int seg_len,
                                                 long long words, uint32_t* __restrict__ sumA, uint32_t* __restrict__ sumB,
                                                 uint32_t* __restrict__ ever_io) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= words) return;
  const int k0 = blockIdx.y * seg_len, k1 = min(m, k0 + seg_len);
  uint32_t A = 0u, B = 0u, E = 0u;
  if (disp) {
#pragma unroll 8
    for (int k = k0; k < k1; ++k) {
      const uint32_t v = __ldg(vis + (long long)k * words + c), d = __ldg(disp + (long long)k * words + c);
      A |= d;
      B = (B & ~d) | v;
      E |= v;
    }
    sumA[(long long)blockIdx.y * words + c] = A;
    sumB[(long long)blockIdx.y * words + c] = B;
  } else {
#pragma unroll 8
    for (int k = k0; k < k1; ++k) E |= __ldg(vis + (long long)k * words + c);
  }
  if (E) atomicOr(ever_io + c, E);
}

__global__ void __launch_bounds__(128) k_cov_apply(const uint32_t* __restrict__ vis, uint32_t* disp_group, int m, int seg_len, long long words,
                                                   const uint32_t* __restrict__ sumA, const uint32_t* __restrict__ sumB,
                                                   const uint32_t* __restrict__ cs_in, uint32_t* __restrict__ cs_out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= words) return;
  const int seg = blockIdx.y;
  const int k0 = seg * seg_len, k1 = min(m, k0 + seg_len);
  uint32_t cs = __ldg(cs_in + c);
  for (int s = 0; s < seg; ++s) cs = (cs & ~__ldg(sumA + (long long)s * words + c)) | __ldg(sumB + (long long)s * words + c);
  int k = k0;
  for (; k + 4 <= k1; k += 4) {
    uint32_t v[4], d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { v[u] = __ldg(vis + (long long)(k + u) * words + c); d[u] = disp_group[(long long)(k + u) * words + c]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      cs &= ~d[u];
      disp_group[(long long)(k + u) * words + c] = v[u] & ~cs;
      cs |= v[u];
    }
  }
  for (; k < k1; ++k) {
    const uint32_t v = __ldg(vis + (long long)k * words + c), d = disp_group[(long long)k * words + c];
    cs &= ~d;
    disp_group[(long long)k * words + c] = v & ~cs;
    cs |= v;
  }
  if (seg == (int)gridDim.y - 1) cs_out[c] = cs;
}

}  // namespace fcvm

// ------------------------------------------------------------------------------------------------
using namespace fcvm;

struct fcvm_cov_handle {
  fcvm_params prm;
  fcvm_camera camp;
  Camera cam;            // FOV normals (PerceptionUtils)
  CovConst cc;
  int threads = 1;
  int64_t P = 0, W = 0;
  int nchunks = 0;
  int slots = 0;
  DevBuf<float4> d_points;
  DevBuf<float> d_box;
  DevBuf<CovPose> d_poses;
  PinBuf<CovPose> h_poses;
  DevBuf<uint32_t> d_vis, d_disp, d_cs, d_cs_in, d_sumA, d_sumB, d_ever, d_indices;
  DevBuf<unsigned long long> d_zdist, d_counters;
  DevBuf<int> d_zidx, d_touched, d_next, d_count;
  DevBuf<long long> d_offsets;
  PinBuf<long long> h_offsets;
  PinBuf<uint32_t> h_indices, h_words;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm = nullptr;
  bool device_ok = false, cloud_ok = false;
  double last_kernel_ms = 0.0, last_scan_ms = 0.0;
  std::vector<int32_t> last_ids;               // CSR ids of the most recent pinhole / evaluate call (fcvm_cov_fetch_ids)
  int64_t counters[6] = {0, 0, 0, 0, 0, 0};   // pair tests, projections, out of image, kernels, chunks scanned, slow paths
  std::unique_ptr<WorkerPool> pool;            // host threads for the per-pose records (spawning threads per call cost ~0.3 ms of a 1.9 ms call)
};

static int cov_ensure_device(fcvm_cov_handle* h) {
  if (h->device_ok) return FCVM_OK;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(FCVM_ENODEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                    " (this library has no CPU fallback)");
  CU(cudaSetDevice(h->prm.device));
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&h->ev0));
  CU(cudaEventCreate(&h->ev1));
  CU(cudaEventCreate(&h->evm));
  CU(h->d_counters.reserve(4));
  CU(cudaMemset(h->d_counters.p, 0, 4 * sizeof(unsigned long long)));
  CU(h->d_next.reserve(1));
  CU(cudaFuncSetAttribute(k_cov_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cov_smem_bytes()));   // 72 KB: opt-in, per device
  // z-buffer slots: one per resident CTA
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->prm.device);
  const long long npix = (long long)h->cc.xPixel * h->cc.yPixel;
  h->slots = sms * kCovCtasPerSm;
  while (h->slots > 1 && (long long)h->slots * npix * 16 > (8ll << 30)) h->slots /= 2;
  CU(h->d_zdist.reserve((size_t)h->slots * npix));
  CU(h->d_zidx.reserve((size_t)h->slots * npix));
  CU(h->d_touched.reserve((size_t)h->slots * npix));
  CU(cudaMemset(h->d_zdist.p, 0xFF, (size_t)h->slots * npix * sizeof(unsigned long long)));
  h->device_ok = true;
  return FCVM_OK;
}

// One batch of poses [k0, k0 + mb): rows on the device; sweep; the requested rows as CSR appended to the outputs.
static int cov_batch(fcvm_cov_handle* h, const double* pose5, int64_t mb, bool evaluate, int64_t* offsets, int32_t* ids, int64_t cap, int64_t& total,
                     float& kernel_ms, float& scan_ms) {
  const int64_t W = h->W;
  CU(h->h_poses.reserve((size_t)mb));
  if (!h->pool) h->pool.reset(new WorkerPool(h->threads));
  h->pool->run(mb, 256, [&](int64_t i) {
    const double* p = pose5 + 5 * i;
    CovPose& cp = h->h_poses.p[i];
    h->cam.fill(cp.rec, p[0], p[1], p[2], p[3], p[4]);
    camera_pose_matrix(p[3], p[4], cp.M);
    cp.pad[0] = cp.pad[1] = cp.pad[2] = 0.0;
  });
  CU(h->d_poses.reserve((size_t)mb));
  CU(h->d_vis.reserve((size_t)mb * W));
  if (evaluate) CU(h->d_disp.reserve((size_t)mb * W));
  CU(h->d_count.reserve((size_t)mb));
  CU(h->d_offsets.reserve((size_t)mb + 1));
  cudaStream_t s = h->stream;
  CU(cudaMemcpyAsync(h->d_poses.p, h->h_poses.p, (size_t)mb * sizeof(CovPose), cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(h->d_vis.p, 0, (size_t)mb * W * sizeof(uint32_t), s));
  if (evaluate) CU(cudaMemsetAsync(h->d_disp.p, 0, (size_t)mb * W * sizeof(uint32_t), s));
  CU(cudaMemsetAsync(h->d_next.p, 0, sizeof(int), s));

  CovArgs a;
  a.fov.max_dist = h->prm.max_dist;
  const double md2 = h->prm.max_dist * h->prm.max_dist;
  a.fov.md2_lo = md2 * (1.0 - 1e-12);
  a.fov.md2_hi = md2 * (1.0 + 1e-12);
  a.fov.plane_eps = 1e-10 * std::max(h->prm.max_dist, 1.0);
  a.fov.visible_range = h->prm.visible_range;
  a.fov.vr2_lo = a.fov.vr2_hi = h->prm.visible_range * h->prm.visible_range;      // (E1 only; unused here)
  a.cam = h->cc;
  a.points = h->d_points.p;
  a.P = h->P;
  a.chunk_box = h->d_box.p;
  a.nchunks = h->nchunks;
  a.poses = h->d_poses.p;
  a.m = (int)mb;
  a.vis_rows = h->d_vis.p;
  a.disp_rows = evaluate ? h->d_disp.p : nullptr;
  a.words = W;
  a.zdist = h->d_zdist.p;
  a.zidx = h->d_zidx.p;
  a.touched = h->d_touched.p;
  a.npix = (long long)h->cc.xPixel * h->cc.yPixel;
  a.next_pose = h->d_next.p;
  a.counters = h->d_counters.p;
  const int grid = (int)std::min<int64_t>(mb, h->slots);
  CU(cudaEventRecord(h->ev0, s));
  k_cov_scan<<<grid, kCovThreads, cov_smem_bytes(), s>>>(a);
  CU(cudaGetLastError());
  CU(cudaEventRecord(h->evm, s));
  {
    // segments: enough (segment, column) threads to fill the chip, at least 64 rows each
    int nseg = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)128, (150000 + W - 1) / W, (mb + 63) / 64}));
    const int seg_len = (int)((mb + nseg - 1) / nseg);
    nseg = (int)((mb + seg_len - 1) / seg_len);
    const dim3 g((unsigned)((W + 127) / 128), (unsigned)nseg);
    if (evaluate) {
      CU(h->d_sumA.reserve((size_t)nseg * W));
      CU(h->d_sumB.reserve((size_t)nseg * W));
    }
    k_cov_sum<<<g, 128, 0, s>>>(h->d_vis.p, evaluate ? h->d_disp.p : nullptr, (int)mb, seg_len, W, h->d_sumA.p, h->d_sumB.p, h->d_ever.p);
    CU(cudaGetLastError());
    if (evaluate) {
      // cs_io is read by every segment's threads and written by the last segment's: double-buffer it
      CU(cudaMemcpyAsync(h->d_cs_in.p, h->d_cs.p, (size_t)W * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
      k_cov_apply<<<g, 128, 0, s>>>(h->d_vis.p, h->d_disp.p, (int)mb, seg_len, W, h->d_sumA.p, h->d_sumB.p, h->d_cs_in.p, h->d_cs.p);
      CU(cudaGetLastError());
      h->counters[3] += 1;
    }
  }
  const uint32_t* out_rows = evaluate ? h->d_disp.p : h->d_vis.p;
  const bool want_csr = offsets || ids;          // the path-coverage loop only needs the union
  if (want_csr) {
    CU(launch_popc_rows(out_rows, mb, W, h->d_count.p, s));
    CU(launch_exscan(h->d_count.p, mb, h->d_offsets.p, s));
    h->counters[3] += 2;
  }
  CU(cudaEventRecord(h->ev1, s));
  h->counters[3] += 2;
  long long ct = 0;
  if (want_csr) {
    CU(h->h_offsets.reserve((size_t)mb + 1));
    CU(cudaMemcpyAsync(h->h_offsets.p, h->d_offsets.p, (size_t)(mb + 1) * sizeof(long long), cudaMemcpyDeviceToHost, s));
  }
  CU(cudaStreamSynchronize(s));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  kernel_ms += ms;
  CU(cudaEventElapsedTime(&ms, h->ev0, h->evm));
  scan_ms += ms;
  if (want_csr) {
    ct = h->h_offsets.p[mb];
    if (offsets)
      for (int64_t i = 0; i < mb; ++i) offsets[i + 1] = total + h->h_offsets.p[i + 1];
  }
  if (ct > 0) {
    CU(h->d_indices.reserve((size_t)ct));
    CU(h->h_indices.reserve((size_t)ct));
    CU(launch_compact_rows(out_rows, mb, W, h->d_offsets.p, h->d_indices.p, s));
    h->counters[3] += 1;
    CU(cudaMemcpyAsync(h->h_indices.p, h->d_indices.p, (size_t)ct * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (ids && total < cap) std::memcpy(ids + total, h->h_indices.p, (size_t)std::min<long long>(ct, cap - total) * sizeof(int32_t));
    h->last_ids.insert(h->last_ids.end(), (const int32_t*)h->h_indices.p, (const int32_t*)h->h_indices.p + ct);
  }
  total += ct;
  return FCVM_OK;
}

static int64_t cov_run(fcvm_cov_handle* h, const double* pose5, int64_t m, bool evaluate, int64_t* offsets, int32_t* ids, int64_t cap, uint8_t* mark,
                       double* rate) {
  if (!h || (m > 0 && !pose5) || m < 0) return fail(FCVM_EINVAL, "bad arguments");
  if (!h->cloud_ok) return fail(FCVM_ENOTREADY, "set the cloud first");
  int rc = cov_ensure_device(h);
  if (rc) return rc;
  const int64_t W = h->W;
  CU(h->d_cs.reserve((size_t)W));
  CU(h->d_cs_in.reserve((size_t)W));
  CU(h->d_ever.reserve((size_t)W));
  CU(cudaMemsetAsync(h->d_cs.p, 0, (size_t)W * sizeof(uint32_t), h->stream));
  CU(cudaMemsetAsync(h->d_ever.p, 0, (size_t)W * sizeof(uint32_t), h->stream));
  int64_t total = 0;
  float kernel_ms = 0.f, scan_ms = 0.f;
  h->last_ids.clear();
  if (offsets) offsets[0] = 0;
  const int64_t per = std::max<int64_t>(1, std::min<int64_t>((int64_t)(2ll << 30) / (W * 4), INT_MAX / 2));
  for (int64_t k0 = 0; k0 < m; k0 += per) {
    const int64_t mb = std::min(per, m - k0);
    rc = cov_batch(h, pose5 + 5 * k0, mb, evaluate, offsets ? offsets + k0 : nullptr, ids, cap, total, kernel_ms, scan_ms);
    if (rc) return rc;
  }
  h->last_kernel_ms = kernel_ms;
  h->last_scan_ms = scan_ms;
  h->counters[0] += m * h->P;
  unsigned long long c[4];
  CU(cudaMemcpyAsync(c, h->d_counters.p, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemsetAsync(h->d_counters.p, 0, sizeof(c), h->stream));
  CU(h->h_words.reserve((size_t)W));
  CU(cudaMemcpyAsync(h->h_words.p, h->d_ever.p, (size_t)W * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->counters[1] += (int64_t)c[0];
  h->counters[2] += (int64_t)c[1];
  h->counters[4] += (int64_t)c[2];
  h->counters[5] += (int64_t)c[3];
  int64_t truenum = 0;
  for (int64_t i = 0; i < h->P; ++i) {
    const uint8_t b = (uint8_t)((h->h_words.p[i >> 5] >> (i & 31)) & 1u);
    truenum += b;
    if (mark) mark[i] = b;
  }
  if (rate) *rate = (double)truenum / (double)h->P;     // hcplanner.cpp:1113-1117
  return total;
}

extern "C" {

fcvm_cov_t* fcvm_cov_create(const fcvm_params* p, const fcvm_camera* cam) {
  if (!p || !cam) { fail(FCVM_EINVAL, "null params"); return nullptr; }
  const int xPixel = 2 * (int)cam->cx, yPixel = 2 * (int)cam->cy;     // hcplanner.cpp:1039, :1131
  if (xPixel <= 0 || yPixel <= 0 || (long long)xPixel * yPixel > (1ll << 28)) { fail(FCVM_EINVAL, "bad image size"); return nullptr; }
  fcvm_cov_handle* h = new (std::nothrow) fcvm_cov_handle();
  if (!h) { fail(FCVM_EINTERNAL, "out of memory"); return nullptr; }
  h->prm = *p;
  h->camp = *cam;
  h->cam.init(*p);
  const double xRange = cam->fx * std::tan(0.5 * cam->fov_w * M_PI / 180.0) / 1000.0;   // :1040
  const double yRange = cam->fy * std::tan(0.5 * cam->fov_h * M_PI / 180.0) / 1000.0;   // :1041
  h->cc.fx = cam->fx; h->cc.fy = cam->fy;
  h->cc.ldc0 = -xRange; h->cc.ldc1 = -yRange;
  h->cc.xPixel = xPixel; h->cc.yPixel = yPixel;
  h->cc.xRes = 2 * xRange / xPixel; h->cc.yRes = 2 * yRange / yPixel;                   // :1044
  const int t = p->host_threads > 0 ? p->host_threads : (int)std::thread::hardware_concurrency();
  h->threads = std::max(1, std::min(t, 64));
  return h;
}

void fcvm_cov_destroy(fcvm_cov_t* h) {
  if (!h) return;
  if (h->device_ok) {
    cudaSetDevice(h->prm.device);
    cudaStreamSynchronize(h->stream);
    h->d_points.free(); h->d_box.free(); h->d_poses.free(); h->h_poses.free();
    h->d_vis.free(); h->d_disp.free(); h->d_cs.free(); h->d_cs_in.free(); h->d_sumA.free(); h->d_sumB.free(); h->d_ever.free(); h->d_indices.free();
    h->d_zdist.free(); h->d_counters.free(); h->d_zidx.free(); h->d_touched.free(); h->d_next.free(); h->d_count.free();
    h->d_offsets.free(); h->h_offsets.free(); h->h_indices.free(); h->h_words.free();
    cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->evm);
    cudaStreamDestroy(h->stream);
  }
  delete h;
}

int fcvm_cov_set_cloud(fcvm_cov_t* h, const float* xyz, int64_t n) {
  if (!h || !xyz || n <= 0) return fail(FCVM_EINVAL, "bad arguments");
  if (n >= (1ll << 31) - kCovChunk) return fail(FCVM_EINVAL, "cloud too large (the reference indexes points with int)");
  int rc = cov_ensure_device(h);
  if (rc) return rc;
  h->P = n;
  h->W = ((n + 31) / 32 + 3) / 4 * 4;
  h->nchunks = (int)((n + kCovChunk - 1) / kCovChunk);
  std::vector<float4> pts((size_t)n);
  std::vector<float> box((size_t)h->nchunks * 6);
  parallel_for(h->nchunks, h->threads, [&](int64_t c) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int64_t i = c * kCovChunk; i < std::min<int64_t>(n, (c + 1) * kCovChunk); ++i) {
      pts[(size_t)i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f);
      for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], xyz[3 * i + d]); mx[d] = std::max(mx[d], xyz[3 * i + d]); }
    }
    for (int d = 0; d < 3; ++d) { box[(size_t)c * 6 + d] = mn[d]; box[(size_t)c * 6 + 3 + d] = mx[d]; }
  });
  CU(h->d_points.reserve((size_t)n));
  CU(h->d_box.reserve(box.size()));
  CU(cudaMemcpy(h->d_points.p, pts.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_box.p, box.data(), box.size() * sizeof(float), cudaMemcpyHostToDevice));
  h->cloud_ok = true;
  return FCVM_OK;
}

int64_t fcvm_cov_pinhole(fcvm_cov_t* h, const double* pose5, int64_t m, int64_t* offsets, int32_t* ids, int64_t cap, uint8_t* covered) {
  return cov_run(h, pose5, m, false, offsets, ids, cap, covered, nullptr);
}

int64_t fcvm_cov_evaluate(fcvm_cov_t* h, const double* pose5, int64_t m, double* rate, int64_t* group_offsets, int32_t* group_ids, int64_t cap,
                          uint8_t* ever) {
  return cov_run(h, pose5, m, true, group_offsets, group_ids, cap, ever, rate);
}

int64_t fcvm_cov_fetch_ids(fcvm_cov_t* h, int32_t* ids, int64_t cap) {
  if (!h || (cap > 0 && !ids)) return fail(FCVM_EINVAL, "bad arguments");
  const int64_t n = (int64_t)h->last_ids.size();
  if (n > 0 && cap > 0) std::memcpy(ids, h->last_ids.data(), (size_t)std::min(n, cap) * sizeof(int32_t));
  return n;
}

int32_t fcvm_cov_chunk_points(void) { return kCovChunk; }
int fcvm_cov_counters(fcvm_cov_t* h, int64_t* out6) {
  if (!h || !out6) return fail(FCVM_EINVAL, "bad arguments");
  for (int k = 0; k < 6; ++k) out6[k] = h->counters[k];
  return FCVM_OK;
}

double fcvm_cov_last_kernel_ms(fcvm_cov_t* h) { return h ? h->last_kernel_ms : 0.0; }
double fcvm_cov_last_scan_ms(fcvm_cov_t* h) { return h ? h->last_scan_ms : 0.0; }

}  // extern "C"
