// fcvm_kernels.cu — sm_100a kernels of the viewpoint visibility / coverage path.
//
//   k_eval<STAGE>   candidate-viewpoint x surface-point visibility (FOV + range + BiRC [+ offset
//                   walk]) -> per-viewpoint coverage bitsets.  The batched form of the loops at
//                   viewpoint_manager.cpp:358-425 (E1), 457-498 (E2), 822-875 (E3), 546-586 (E4).
//   k_popc_rows     contribute_voxel_num per viewpoint = popcount of its bitset row.
//   k_own_sweep     the ownership rule (viewpoint_manager.cpp:500-528, 589-617) over bitset rows.
//
// Layout of one k_eval CTA: a tile of model points is staged ONCE into shared memory with a 1-D
// TMA bulk copy (cp.async.bulk + mbarrier), then each warp pulls viewpoints from a CTA-local
// counter and tests its 32 lanes' points against the viewpoint's four FOV planes and range.
// Pairs that pass are compacted by ballot into a per-warp ray queue; whenever 32 rays are queued
// the whole warp marches them (one ray per lane, both half-rays in lock step like the
// reference), so the expensive divergent phase always runs with full warps.  Visible bits are
// OR-ed into the viewpoint's bitset row in HBM.  No tensor cores: nothing here is a contraction.
#include "fcvm_device.cuh"
#include "fcvm_kernels.h"

namespace fcvm {

// ------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy wrappers (PTX; SASS: UBLKCP + SYNCS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// k_eval
// ------------------------------------------------------------------------------------------
constexpr int kWarps = EVAL_WARPS;
constexpr int kQueue = 64;  // ray-queue entries per warp

struct QueueEntry {
  uint32_t vp;   // viewpoint index inside the batch
  uint32_t pt;   // point index inside the tile
};

template <int STAGE>
__device__ __forceinline__ bool cast_one(const GridDesc& g, const FovConst& fc, const VpRec* __restrict__ recs, uint32_t v,
                                         float4 q, RayStats& st) {
  const double vx = __ldg(&recs[v].px), vy = __ldg(&recs[v].py), vz = __ldg(&recs[v].pz);
  const double qx = (double)q.x, qy = (double)q.y, qz = (double)q.z;
  if (STAGE == 1) {
    // E1 also requires ray_length <= visible_range (viewpoint_manager.cpp:402-403), tested by the
    // reference after the cast; the ray is marched either way so the lookup count matches.
    const double dx = qx - vx, dy = qy - vy, dz = qz - vz;
    const double len = sqrt((dx * dx + dy * dy) + dz * dz);
    const bool vis = birc(g, vx, vy, vz, qx, qy, qz, true, st);
    return vis && (len <= fc.visible_range);
  } else if (STAGE == 3) {
    return birc(g, vx, vy, vz, qx, qy, qz, true, st);
  } else {
    double tx, ty, tz;
    offset_target(g, qx, qy, qz, vx, vy, vz, tx, ty, tz, st);
    return birc(g, vx, vy, vz, tx, ty, tz, STAGE == 2, st);
  }
}

template <int STAGE>
__global__ void __launch_bounds__(kWarps * 32) k_eval(const __grid_constant__ EvalArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4* tile = reinterpret_cast<float4*>(smem);
  QueueEntry* queues = reinterpret_cast<QueueEntry*>(smem + (size_t)a.tile_pts * sizeof(float4));
  __shared__ __align__(8) uint64_t bar;
  __shared__ int next_vp;
  __shared__ float red[kWarps][6];
  __shared__ double aabb[6];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long tile_base = (long long)blockIdx.x * a.tile_pts;
  const int npts = (int)min((long long)a.tile_pts, a.P - tile_base);
  // viewpoint chunk of this CTA
  const long long v_lo = (long long)blockIdx.y * a.vp_per_cta;
  const long long v_hi = min(v_lo + (long long)a.vp_per_cta, a.V);

  // ---- stage the point tile with one TMA bulk copy --------------------------------------
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    next_vp = 0;
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)npts * (uint32_t)sizeof(float4);
    mbar_expect_tx(&bar, bytes);
    tma_bulk_g2s(tile, a.points + tile_base, bytes, &bar);
  }
  mbar_wait(&bar, 0);

  // ---- tile bounding box (for the conservative whole-tile range cull) ---------------------
  {
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = tid; i < npts; i += kWarps * 32) {
      const float4 q = tile[i];
      mn[0] = fminf(mn[0], q.x); mx[0] = fmaxf(mx[0], q.x);
      mn[1] = fminf(mn[1], q.y); mx[1] = fmaxf(mx[1], q.y);
      mn[2] = fminf(mn[2], q.z); mx[2] = fmaxf(mx[2], q.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
        mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
      }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { red[warp][c] = mn[c]; red[warp][3 + c] = mx[c]; }
    }
    __syncthreads();
    if (tid < 6) {
      float r = red[0][tid];
      for (int w = 1; w < kWarps; ++w) r = tid < 3 ? fminf(r, red[w][tid]) : fmaxf(r, red[w][tid]);
      aabb[tid] = (double)r;
    }
    __syncthreads();
  }

  const GridDesc& g = a.grid;
  const FovConst& fc = a.fov;
  QueueEntry* q = queues + warp * kQueue;
  int qn = 0;                       // warp-uniform queue length
  RayStats st = {0u, 0u, 0u};
  unsigned int nrays = 0;
  unsigned long long npairs = 0;
  const long long W = a.words;

  // march one queued ray per lane (lanes >= cnt idle), OR visible bits into the bitset rows
  auto flush = [&](int cnt) {
    const int e = qn - cnt + lane;
    if (lane < cnt) {
      const QueueEntry qe = q[e];
      const float4 pt = tile[qe.pt];
      nrays++;
      const bool vis = cast_one<STAGE>(g, fc, a.recs, qe.vp, pt, st);
      if (vis) {
        const long long gp = tile_base + qe.pt;
        atomicOr(a.rows + (long long)qe.vp * W + (gp >> 5), 1u << (gp & 31));
      }
    }
    qn -= cnt;
    __syncwarp();
  };

  for (;;) {
    int vi = 0;
    if (lane == 0) vi = atomicAdd(&next_vp, 1);
    vi = __shfl_sync(0xffffffffu, vi, 0);
    const long long v = v_lo + vi;
    if (v >= v_hi) break;

    // viewpoint record -> registers (uniform loads)
    VpRec rec;
    {
      const double2* src = reinterpret_cast<const double2*>(a.recs + v);
      double2* dst = reinterpret_cast<double2*>(&rec);
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[k] = __ldg(src + k);
    }
    // whole-tile cull: squared distance from the viewpoint to the tile's bounding box.  Every point
    // of a culled tile fails `dir.norm() > max_dist` in the reference, so skipping is exact.
    {
      const double cx = fmax(fmax(aabb[0] - rec.px, rec.px - aabb[3]), 0.0);
      const double cy = fmax(fmax(aabb[1] - rec.py, rec.py - aabb[4]), 0.0);
      const double cz = fmax(fmax(aabb[2] - rec.pz, rec.pz - aabb[5]), 0.0);
      if ((cx * cx + cy * cy) + cz * cz > fc.md2_hi * 1.000001) {
        if (STAGE != 2) npairs += (lane == 0) ? (unsigned long long)npts : 0ull;
        continue;
      }
    }
    const uint32_t* rrow = (STAGE == 2) ? a.restrict_rows + v * W + (tile_base >> 5) : nullptr;
    for (int c = 0; c < npts; c += 32) {
      const int i = c + lane;
      bool valid = i < npts;
      if (STAGE == 2) {
        const uint32_t rw = __ldg(rrow + (c >> 5));   // only points E1 found visible are re-tested
        if (rw == 0u) continue;
        valid = valid && ((rw >> lane) & 1u);
      }
      bool pass = false;
      if (valid) {
        const float4 pt = tile[i];
        pass = inside_fov(rec, fc, (double)pt.x, (double)pt.y, (double)pt.z);
        npairs++;
      }
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (m == 0u) continue;
      if (pass) {
        const int pos = qn + __popc(m & ((1u << lane) - 1u));
        q[pos].vp = (uint32_t)v;
        q[pos].pt = (uint32_t)i;
      }
      qn += __popc(m);
      __syncwarp();
      if (qn >= 32) flush(32);
    }
  }
  if (qn > 0) flush(qn);   // drain (qn < 32 here)

  // ---- counters -------------------------------------------------------------------------
  if (a.counters) {
    unsigned long long c0 = npairs, c1 = nrays, c2 = st.lookups, c3 = st.walk, c4 = st.trips;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o);
      c3 += __shfl_xor_sync(0xffffffffu, c3, o);
      c4 += __shfl_xor_sync(0xffffffffu, c4, o);
    }
    if (lane == 0) {
      if (c0) atomicAdd(a.counters + 0, c0);
      if (c1) atomicAdd(a.counters + 1, c1);
      if (c2) atomicAdd(a.counters + 2, c2);
      if (c3) atomicAdd(a.counters + 3, c3);
      if (c4) atomicAdd(a.counters + 4, c4);
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_popc_rows: count[r] = popcount(row r).  One warp per row, 128-bit loads.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_popc_rows(const uint32_t* __restrict__ rows, long long n, long long words, int* __restrict__ count) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const int lane = threadIdx.x & 31;
  const uint32_t* row = rows + r * words;
  int c = 0;
  const long long w4 = words >> 2;
  const uint4* row4 = reinterpret_cast<const uint4*>(row);   // rows are 16-byte aligned (words % 4 == 0)
  for (long long i = lane; i < w4; i += 32) {
    const uint4 x = __ldg(row4 + i);
    c += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) count[r] = c;
}

// ------------------------------------------------------------------------------------------
// k_exscan: offsets[0..n] = exclusive prefix sum of count[0..n) (one block; n is a batch size).
// k_compact_rows: bitset rows -> ascending point-index lists (CSR).  The visible sets are 1-4 %
// dense, so the host-side consumers (the ordered pitch / yaw sums) read 4 bytes per visible ray
// instead of P/8 bytes per viewpoint.  One warp per row; lane order = word order keeps indices
// ascending.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_exscan(const int* __restrict__ count, long long n, long long* __restrict__ offsets) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const long long per = (n + 1023) / 1024;
  const long long lo = min(n, per * t), hi = min(n, per * (t + 1));
  long long s = 0;
  for (long long i = lo; i < hi; ++i) s += count[i];
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    long long v = t >= o ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  long long run = t ? part[t - 1] : 0;
  for (long long i = lo; i < hi; ++i) { offsets[i] = run; run += count[i]; }
  if (t == 1023) offsets[n] = part[1023];
}

__global__ void __launch_bounds__(256) k_compact_rows(const uint32_t* __restrict__ rows, long long n, long long words,
                                                      const long long* __restrict__ offsets, uint32_t* __restrict__ indices) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const int lane = threadIdx.x & 31;
  const uint32_t* row = rows + r * words;
  long long base = offsets[r];
  for (long long w0 = 0; w0 < words; w0 += 32) {
    const long long wi = w0 + lane;
    uint32_t w = wi < words ? __ldg(row + wi) : 0u;
    if (__ballot_sync(0xffffffffu, w != 0u) == 0u) continue;
    const int c = __popc(w);
    int pre = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += v;
    }
    long long pos = base + (pre - c);
    while (w) {
      const int b = __ffs(w) - 1;
      w &= w - 1;
      indices[pos++] = (uint32_t)(wi * 32 + b);
    }
    base += __shfl_sync(0xffffffffu, pre, 31);
  }
}

// ------------------------------------------------------------------------------------------
// k_own_sweep: the ownership rule of getPose / updateViewpointInfo
// (viewpoint_manager.cpp:500-528, 589-617), for a whole stage at once.
//   for each row in processing order (sched[t] = {row, vp_id, contribute_voxel_num > 0}):
//     for each visible point j:  uncovered -> take;   covered and contri > cover_contrib[j] and
//     current owner id >= last_vps_num -> take over (owner's count -1);   take: count[vp] +1.
// One warp per 32-point word column; the lane IS the point, so the per-point state lives in
// registers for the whole sweep; the row words of a column are warp-uniform loads.  Counts are
// accumulated with ballot/popc per warp and one atomic per (warp, row).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_own_sweep(const uint32_t* __restrict__ rows, long long words, const int4* __restrict__ sched,
                                                   int nsched, int last_vps_num, long long P, int* __restrict__ cover_contrib,
                                                   int* __restrict__ contrib_id, int* __restrict__ voxcount) {
  const long long col = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (col * 32 >= P) return;
  const int lane = threadIdx.x & 31;
  const long long j = col * 32 + lane;
  const bool inb = j < P;
  int contrib = inb ? cover_contrib[j] : 0;
  int owner = inb ? contrib_id[j] : -1;
  for (int t = 0; t < nsched; ++t) {
    const int4 s = __ldg(sched + t);          // x = row, y = vp id, z = contri
    const uint32_t w = __ldg(rows + (long long)s.x * words + col);
    if (w == 0u) continue;
    const bool seen = inb && ((w >> lane) & 1u);
    bool take = false;
    if (seen) {
      if (owner < 0) {
        take = true;
      } else if (s.z > contrib && owner >= last_vps_num) {
        atomicSub(voxcount + owner, 1);
        take = true;
      }
      if (take) { contrib = s.z; owner = s.y; }
    }
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (m != 0u && lane == 0) atomicAdd(voxcount + s.y, __popc(m));
  }
  if (inb) { cover_contrib[j] = contrib; contrib_id[j] = owner; }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <int STAGE>
static cudaError_t launch_eval_t(const EvalArgs& a, dim3 grid, size_t smem, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k_eval<STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eval_smem_bytes(EVAL_MAX_TILE));
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  k_eval<STAGE><<<grid, kWarps * 32, smem, s>>>(a);
  return cudaGetLastError();
}

size_t eval_smem_bytes(int tile_pts) { return (size_t)tile_pts * sizeof(float4) + (size_t)kWarps * kQueue * sizeof(QueueEntry); }

cudaError_t launch_eval(int stage, const EvalArgs& a, cudaStream_t s) {
  if (a.V <= 0 || a.P <= 0) return cudaSuccess;
  const long long tiles = (a.P + a.tile_pts - 1) / a.tile_pts;
  const long long chunks = (a.V + a.vp_per_cta - 1) / a.vp_per_cta;
  if (chunks > 65535) return cudaErrorInvalidConfiguration;
  dim3 grid((unsigned)tiles, (unsigned)chunks);
  const size_t smem = eval_smem_bytes(a.tile_pts);
  switch (stage) {
    case 1: return launch_eval_t<1>(a, grid, smem, s);
    case 2: return launch_eval_t<2>(a, grid, smem, s);
    case 3: return launch_eval_t<3>(a, grid, smem, s);
    case 4: return launch_eval_t<4>(a, grid, smem, s);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_popc_rows(const uint32_t* rows, long long n, long long words, int* count, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const int wpb = 8;
  k_popc_rows<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, s>>>(rows, n, words, count);
  return cudaGetLastError();
}

cudaError_t launch_exscan(const int* count, long long n, long long* offsets, cudaStream_t s) {
  k_exscan<<<1, 1024, 0, s>>>(count, n, offsets);
  return cudaGetLastError();
}

cudaError_t launch_compact_rows(const uint32_t* rows, long long n, long long words, const long long* offsets, uint32_t* indices, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const int wpb = 8;
  k_compact_rows<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, s>>>(rows, n, words, offsets, indices);
  return cudaGetLastError();
}

cudaError_t launch_own_sweep(const uint32_t* rows, long long words, const int4* sched, int nsched, int last_vps_num, long long P,
                             int* cover_contrib, int* contrib_id, int* voxcount, cudaStream_t s) {
  if (nsched <= 0 || P <= 0) return cudaSuccess;
  const long long cols = (P + 31) / 32;
  const int wpb = 4;
  k_own_sweep<<<(unsigned)((cols + wpb - 1) / wpb), wpb * 32, 0, s>>>(rows, words, sched, nsched, last_vps_num, P, cover_contrib, contrib_id, voxcount);
  return cudaGetLastError();
}

}  // namespace fcvm
