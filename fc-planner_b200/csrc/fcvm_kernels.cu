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
// Pairs that pass are compacted by ballot into a per-warp entry queue.  32 entries at a time are
// turned into prepared rays (biInput at full warp width) and pushed onto a per-warp queue of ray
// continuations in shared memory; the march loop then runs one ray per lane, both half-rays in
// lock step like the reference, and a lane whose ray ends (blocked early, or short) pulls the next
// continuation immediately, so the divergent phase keeps its lanes busy although ray lengths vary
// by two orders of magnitude.  Visible bits are OR-ed into the viewpoint's bitset row in HBM.
// No tensor cores: nothing here is a contraction.
#include <atomic>

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
// FCVM_DEBUG_BOUNDS builds (libfcvm_dbg.so, tests/test_gpu_bounds.py) check every data-dependent
// index of k_eval - queue slots, tile offsets, bitset words, brick indices - and count violations in
// counter slot 5 instead of performing the access.  compute-sanitizer is closed on this GPU pool,
// so this is the memory-safety evidence.  Release builds compile the checks away.
#ifndef EVAL_ASM_STEP
#define EVAL_ASM_STEP 1
#endif
#ifndef EVAL_STATS
#define EVAL_STATS 0          // 1 / 2 / 3: measurement builds that reuse the walk / trips counters of an E1 pass (profiles/march_stats.py)
#endif
#ifndef EVAL_HANDOFF
#define EVAL_HANDOFF 1        // freshly prepared rays enter the march in registers, not through the shared-memory queue
#endif
#ifndef EVAL_RANGE_SQ
#define EVAL_RANGE_SQ 1         // E1's `ray_length <= visible_range` from the squared length outside a rounding band
#endif
#ifndef EVAL_FAST_WALK
#define EVAL_FAST_WALK 1         // E2 / E4: the surface-offset walk on the fast marcher where that is legal
#endif
#ifndef EVAL_STAGED_FOV
#define EVAL_STAGED_FOV 1
#endif
#ifndef EVAL_BRICK_CHECK
#ifdef FCVM_DEBUG_BOUNDS
#define EVAL_BRICK_CHECK 1
#else
#define EVAL_BRICK_CHECK 0
#endif
#endif
#ifdef FCVM_DEBUG_BOUNDS
#define FCVM_CHECK(cond) ((cond) ? true : (a.counters ? (atomicAdd(a.counters + 5, 1ull), false) : false))
#else
#define FCVM_CHECK(cond) true
#endif
constexpr int kWarps = EVAL_WARPS;
constexpr int kEntryQ = 64;   // (viewpoint, point) pairs that passed the FOV test, per warp
constexpr int kReadyQ = EVAL_READYQ;   // prepared / suspended rays ("continuations"), per warp; >= EVAL_KEEP - 1 + 32
static_assert(EVAL_READYQ >= EVAL_KEEP - 1 + 32, "ready queue too small for a suspended warp plus one prepared batch");
constexpr int kContD = 12;    // doubles per continuation: tMax + tDelta of both half-rays
constexpr int kContI = 11;    // ints per continuation

struct QueueEntry {
  uint32_t vp;   // viewpoint index inside the batch
  uint32_t pt;   // point index inside the tile
};

// One half-ray on the fast path: the voxel is tracked directly as a MAP index (the conversion
// `(int)(x + offset_)` was verified to be a pure shift over the ray's whole extent when the ray was
// prepared), the walk is counted down instead of compared against the midpoint voxel, and the whole
// extent was verified to lie inside the map, so the inner loop has no bounds test.
struct Side {
  int ix, iy, iz, rem;
  int sx, sy, sz;
  double tmx, tmy, tmz, tdx, tdy, tdz;
  // One conditional step: the branch ladder of biNextId (raycast.cpp:496-518: strict '<', z on ties
  // and NaN) evaluated as three predicates so that the updates are predicated, not branched.
  __device__ __forceinline__ void step_if(bool go) {
#if EVAL_ASM_STEP
    // Same ladder, written as predicated PTX: three ordered '<' compares (false on NaN, like the C++ operator), the
    // axis predicates, then one predicated integer add and one predicated FP64 add (add.rn, never fused) per axis.
    // The compiler's own rendering of the C++ below computes all three sums and selects (3 DADD + 6 FSEL + 3 SEL).
    asm("{\n\t"
        ".reg .pred g, xy, xz, yz, px, py, pz;\n\t"
        "setp.ne.s32 g, %13, 0;\n\t"
        "setp.lt.f64 xy, %4, %5;\n\t"
        "setp.lt.f64 xz, %4, %6;\n\t"
        "setp.lt.f64 yz, %5, %6;\n\t"
        "and.pred px, xy, xz;\n\t"
        "not.pred xy, xy;\n\t"
        "and.pred py, xy, yz;\n\t"
        "or.pred pz, px, py;\n\t"
        "not.pred pz, pz;\n\t"
        "and.pred px, px, g;\n\t"
        "and.pred py, py, g;\n\t"
        "and.pred pz, pz, g;\n\t"
        "@px add.s32 %0, %0, %7;\n\t"
        "@px add.rn.f64 %4, %4, %10;\n\t"
        "@py add.s32 %1, %1, %8;\n\t"
        "@py add.rn.f64 %5, %5, %11;\n\t"
        "@pz add.s32 %2, %2, %9;\n\t"
        "@pz add.rn.f64 %6, %6, %12;\n\t"
        "@g add.s32 %3, %3, -1;\n\t"
        "}"
        : "+r"(ix), "+r"(iy), "+r"(iz), "+r"(rem), "+d"(tmx), "+d"(tmy), "+d"(tmz)
        : "r"(sx), "r"(sy), "r"(sz), "d"(tdx), "d"(tdy), "d"(tdz), "r"((int)go));
#else
    const bool xy = tmx < tmy, xz = tmx < tmz, yz = tmy < tmz;
    const bool selx = go && xy && xz;
    const bool sely = go && !xy && yz;
    const bool selz = go && !(xy && xz) && !(!xy && yz);
    if (selx) { ix += sx; tmx += tdx; }
    if (sely) { iy += sy; tmy += tdy; }
    if (selz) { iz += sz; tmz += tdz; }
    rem -= go ? 1 : 0;
#endif
  }
};

#define CONT_DISCARD 1      // prepare only: the call sites make one biNextId call and discard it
#define CONT_INRANGE 2      // E1: ray_length <= visible_range (ANDed into the verdict)
#define CONT_PFREE 4        // every voxel the viewpoint-side half-ray can visit is free (summed-area-table certificate): that
                            // side is not marched; p.ix/iy/iz hold the midpoint voxel, p.rem its remaining biNextId calls
#define CONT_SANE 8         // both half-rays start strictly inside their voxels on every moving axis: each stays in the box spanned
                            // by its current and the midpoint voxel, so "the box between the two fronts is empty" ends the march

struct Cont {
  Side p, n;
  uint32_t vp, pt;
  int flags;
};

__device__ __forceinline__ int pack_dirs(const Side& a, const Side& b, int flags) {
  return (a.sx + 1) | ((a.sy + 1) << 2) | ((a.sz + 1) << 4) | ((b.sx + 1) << 6) | ((b.sy + 1) << 8) | ((b.sz + 1) << 10) | (flags << 12);
}
__device__ __forceinline__ void cont_store(double* qd, int* qi, int slot, const Cont& c) {
  qd[0 * kReadyQ + slot] = c.p.tmx; qd[1 * kReadyQ + slot] = c.p.tmy; qd[2 * kReadyQ + slot] = c.p.tmz;
  qd[3 * kReadyQ + slot] = c.p.tdx; qd[4 * kReadyQ + slot] = c.p.tdy; qd[5 * kReadyQ + slot] = c.p.tdz;
  qd[6 * kReadyQ + slot] = c.n.tmx; qd[7 * kReadyQ + slot] = c.n.tmy; qd[8 * kReadyQ + slot] = c.n.tmz;
  qd[9 * kReadyQ + slot] = c.n.tdx; qd[10 * kReadyQ + slot] = c.n.tdy; qd[11 * kReadyQ + slot] = c.n.tdz;
  qi[0 * kReadyQ + slot] = c.p.ix; qi[1 * kReadyQ + slot] = c.p.iy; qi[2 * kReadyQ + slot] = c.p.iz; qi[3 * kReadyQ + slot] = c.p.rem;
  qi[4 * kReadyQ + slot] = c.n.ix; qi[5 * kReadyQ + slot] = c.n.iy; qi[6 * kReadyQ + slot] = c.n.iz; qi[7 * kReadyQ + slot] = c.n.rem;
  qi[8 * kReadyQ + slot] = pack_dirs(c.p, c.n, c.flags);
  qi[9 * kReadyQ + slot] = (int)c.vp; qi[10 * kReadyQ + slot] = (int)c.pt;
}
__device__ __forceinline__ void cont_load(const double* qd, const int* qi, int slot, Cont& c) {
  c.p.tmx = qd[0 * kReadyQ + slot]; c.p.tmy = qd[1 * kReadyQ + slot]; c.p.tmz = qd[2 * kReadyQ + slot];
  c.p.tdx = qd[3 * kReadyQ + slot]; c.p.tdy = qd[4 * kReadyQ + slot]; c.p.tdz = qd[5 * kReadyQ + slot];
  c.n.tmx = qd[6 * kReadyQ + slot]; c.n.tmy = qd[7 * kReadyQ + slot]; c.n.tmz = qd[8 * kReadyQ + slot];
  c.n.tdx = qd[9 * kReadyQ + slot]; c.n.tdy = qd[10 * kReadyQ + slot]; c.n.tdz = qd[11 * kReadyQ + slot];
  c.p.ix = qi[0 * kReadyQ + slot]; c.p.iy = qi[1 * kReadyQ + slot]; c.p.iz = qi[2 * kReadyQ + slot]; c.p.rem = qi[3 * kReadyQ + slot];
  c.n.ix = qi[4 * kReadyQ + slot]; c.n.iy = qi[5 * kReadyQ + slot]; c.n.iz = qi[6 * kReadyQ + slot]; c.n.rem = qi[7 * kReadyQ + slot];
  const int d = qi[8 * kReadyQ + slot];
  c.p.sx = (d & 3) - 1; c.p.sy = ((d >> 2) & 3) - 1; c.p.sz = ((d >> 4) & 3) - 1;
  c.n.sx = ((d >> 6) & 3) - 1; c.n.sy = ((d >> 8) & 3) - 1; c.n.sz = ((d >> 10) & 3) - 1;
  c.flags = d >> 12;
  c.vp = (uint32_t)qi[9 * kReadyQ + slot]; c.pt = (uint32_t)qi[10 * kReadyQ + slot];
}

// occCheck on the fast path.  No per-axis bounds test: the ray's extent was verified to lie inside the
// map when it was prepared, and a sane marcher never leaves the box spanned by its end voxels.  The
// reference marcher is not always sane (it can run away from a lattice-aligned start, see
// tests/test_oracle_vs_ref.py).  A fast-path half-ray makes at most g.guard_steps steps, and the brick
// array is allocated with guard words on both sides that cover every index reachable in that many
// steps and read "occupied" (fcvm_set_map_cloud): a stray ray never reads outside the allocation and
// is counted as a cap trip when its two halves fail to meet.  The index-checking debug build
// (FCVM_DEBUG_BOUNDS) range-checks every fetch instead.
__device__ __forceinline__ bool occ_fast(const GridDesc& g, int ix, int iy, int iz, int& ckey, unsigned long long& cword) {
#if FCVM_BRICK_LINES
  const int k = ((((ix >> 3) * g.nly + (iy >> 3)) * g.nlz + (iz >> 4)) << 4) | (((ix >> 2) & 1) << 3) | (((iy >> 2) & 1) << 2) | ((iz >> 2) & 3);
#else
  const int k = ((ix >> 2) * g.nby + (iy >> 2)) * g.nbz + (iz >> 2);
#endif
  if (k != ckey) {
    ckey = k;
#if EVAL_BRICK_CHECK
    cword = (unsigned)k < (unsigned)g.nbricks ? __ldg(g.bricks + k) : ~0ull;
#else
    cword = __ldg(g.bricks + k);
#endif
  }
  return ((cword >> brick_bit_index(ix, iy, iz)) & 1ull) == 0ull;
}

// tDelta = step / delta with step = signum(delta) (raycast.cpp:412-434) is 1 / |delta| (NaN for 0 / 0): a table of the
// reciprocals of 0 .. kRcpN - 1, filled once per device with the very division the reference makes
constexpr int kRcpN = EVAL_RCP_TABLE;
__device__ double g_rcp[kRcpN];
__global__ void k_init_rcp() {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < kRcpN) g_rcp[k] = k == 0 ? (0.0 / (double)k) : (1.0 / (double)k);
}
__device__ __forceinline__ double t_delta(int delta) { return __ldg(&g_rcp[abs(delta)]); }   // |delta| < kRcpN on the fast path
__global__ void __launch_bounds__(256) k_ray_ends(const double* __restrict__ xyz, int stride_d, const float4* __restrict__ pts, long long n, double res,
                                                  RayEnd* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (xyz) out[i] = make_ray_end(xyz[i * stride_d], xyz[i * stride_d + 1], xyz[i * stride_d + 2], res);
  else { const float4 q = pts[i]; out[i] = make_ray_end((double)q.x, (double)q.y, (double)q.z, res); }
}

// biInput (raycast.cpp:392-444) for one ray.  Returns true and fills `c` when the ray qualifies for
// the fast path; otherwise the verdict is computed right here with the literal marcher (birc).
// vpe / pte: the two ends divided by the resolution, precomputed per viewpoint / per model point (E2 / E4 aim at the
// offset target instead of the point and divide here).
template <int STAGE>
__device__ __forceinline__ bool prepare_ray(const GridDesc& g, const FovConst& fc, const VpRec* __restrict__ recs, const RayEnd* __restrict__ vp_ends,
                                            const RayEnd* __restrict__ pt_ends, long long gp, uint32_t v, uint32_t pti,
                                            float4 q, Cont& c, bool& slow_verdict, RayStats& st) {
  const double vx = __ldg(&recs[v].px), vy = __ldg(&recs[v].py), vz = __ldg(&recs[v].pz);
  const double qx = (double)q.x, qy = (double)q.y, qz = (double)q.z;
  double bx = qx, by = qy, bz = qz;
  int flags = (STAGE == 4) ? 0 : CONT_DISCARD;
  if (STAGE == 1) {
    // ray_length <= visible_range (viewpoint_manager.cpp:402-403), tested by the reference after the cast
    const double dx = qx - vx, dy = qy - vy, dz = qz - vz;
#if EVAL_RANGE_SQ
    // sqrt is monotone and correctly rounded: outside a 1e-12 relative band around visible_range^2 the squared length decides
    // (NaN fails both compares and takes the literal test, which it fails as in the reference)
    const double n2 = (dx * dx + dy * dy) + dz * dz;
    if (n2 < fc.vr2_lo || (!(n2 > fc.vr2_hi) && sqrt(n2) <= fc.visible_range)) flags |= CONT_INRANGE;
#else
    if (sqrt((dx * dx + dy * dy) + dz * dz) <= fc.visible_range) flags |= CONT_INRANGE;
#endif
  } else {
    flags |= CONT_INRANGE;
  }
  const double res = g.res;
  const double2 ve0 = __ldg(reinterpret_cast<const double2*>(vp_ends + v)), ve1 = __ldg(reinterpret_cast<const double2*>(vp_ends + v) + 1);
  const double sxf = ve0.x, syf = ve0.y, szf = ve1.x;
  int v_inside = (int)ve1.y, n_inside;
  double exf, eyf, ezf;
  if (STAGE == 2 || STAGE == 4) {
    {
      const double2 pe0 = __ldg(reinterpret_cast<const double2*>(pt_ends + gp)), pe1 = __ldg(reinterpret_cast<const double2*>(pt_ends + gp) + 1);
      bool walked = false;
#if EVAL_FAST_WALK
      // The surface-offset walk (point -> viewpoint, stop at the 2nd free voxel) on the fast marcher: map indices tracked
      // directly, steps counted down, unchecked brick fetch.  Legal when both end voxels are inside the map, the map-index
      // conversion is a pure shift between them, every |delta| is below the reciprocal table's size and the point lies strictly
      // inside its voxel on every moving axis: such a marcher arrives after exactly |dx| + |dy| + |dz| steps and never leaves
      // the box of its end voxels (see the certificates below).  Everything else takes the literal walk (offset_target).
      if (((fabs(pe0.x) + fabs(pe0.y)) + fabs(pe1.x)) + ((fabs(sxf) + fabs(syf)) + fabs(szf)) < 1.0e9) {
        const int wx = (int)floor(pe0.x), wy = (int)floor(pe0.y), wz = (int)floor(pe1.x);
        const int ux = (int)floor(sxf), uy = (int)floor(syf), uz = (int)floor(szf);
        const int ddx = ux - wx, ddy = uy - wy, ddz = uz - wz;
        const int wix = to_map_index(wx, g.offx), wiy = to_map_index(wy, g.offy), wiz = to_map_index(wz, g.offz);
        const int uix = to_map_index(ux, g.offx), uiy = to_map_index(uy, g.offy), uiz = to_map_index(uz, g.offz);
        const int in_bits = (int)pe1.y | (ddx == 0 ? 1 : 0) | (ddy == 0 ? 2 : 0) | (ddz == 0 ? 4 : 0);
        if (in_bits == 7 && (unsigned)wix < (unsigned)g.nx && (unsigned)wiy < (unsigned)g.ny && (unsigned)wiz < (unsigned)g.nz &&
            (unsigned)uix < (unsigned)g.nx && (unsigned)uiy < (unsigned)g.ny && (unsigned)uiz < (unsigned)g.nz &&
            uix - wix == ddx && uiy - wiy == ddy && uiz - wiz == ddz && abs(ddx) < kRcpN && abs(ddy) < kRcpN && abs(ddz) < kRcpN) {
          Side w;
          w.ix = wix; w.iy = wiy; w.iz = wiz;
          w.rem = abs(ddx) + abs(ddy) + abs(ddz);
          w.sx = d_signum(ddx); w.sy = d_signum(ddy); w.sz = d_signum(ddz);
          w.tmx = d_intbound_i(pe0.x, ddx, g_rcp); w.tmy = d_intbound_i(pe0.y, ddy, g_rcp); w.tmz = d_intbound_i(pe1.x, ddz, g_rcp);
          w.tdx = t_delta(ddx); w.tdy = t_delta(ddy); w.tdz = t_delta(ddz);
          int key = -1, count = 0, tix, tiy, tiz;
          unsigned long long word = 0ull;
          for (;;) {
            tix = w.ix; tiy = w.iy; tiz = w.iz;          // nextId reports the voxel before stepping
            if (w.rem == 0) break;                        // ... and returns false at the end voxel
            w.step_if(true);
            st.lookups += 1; st.walk += 1;
            if (occ_fast(g, tix, tiy, tiz, key, word) && ++count == 2) break;
          }
          bx = ((double)tix + 0.5) * res + g.orgx;       // indexToPos_hc (sdf_map.h:244-247)
          by = ((double)tiy + 0.5) * res + g.orgy;
          bz = ((double)tiz + 0.5) * res + g.orgz;
          walked = true;
        }
      }
#endif
      if (!walked) offset_target(g, qx, qy, qz, vx, vy, vz, bx, by, bz, st, g_rcp, kRcpN, pe0.x, pe0.y, pe1.x, sxf, syf, szf);
    }
    // the target is a voxel centre: finite, and strictly inside its voxel on every axis unless rounding says otherwise
    exf = div_exact(bx, res, g.res_rcp); eyf = div_exact(by, res, g.res_rcp); ezf = div_exact(bz, res, g.res_rcp);
    const double fx = exf - floor(exf), fy = eyf - floor(eyf), fz = ezf - floor(ezf);
    n_inside = ((fx > 1e-9 && fx < 1.0 - 1e-9) ? 1 : 0) | ((fy > 1e-9 && fy < 1.0 - 1e-9) ? 2 : 0) | ((fz > 1e-9 && fz < 1.0 - 1e-9) ? 4 : 0);
  } else {
    const double2 pe0 = __ldg(reinterpret_cast<const double2*>(pt_ends + gp)), pe1 = __ldg(reinterpret_cast<const double2*>(pt_ends + gp) + 1);
    exf = pe0.x; eyf = pe0.y; ezf = pe1.x; n_inside = (int)pe1.y;
  }
  // inter_ = 0.5 * (start + end) / resolution_: only its floor is used
  const double mxf = div_exact(0.5 * (vx + bx), res, g.res_rcp), myf = div_exact(0.5 * (vy + by), res, g.res_rcp), mzf = div_exact(0.5 * (vz + bz), res, g.res_rcp);
  const int x0 = (int)floor(sxf), y0 = (int)floor(syf), z0 = (int)floor(szf);
  const int mx = (int)floor(mxf), my = (int)floor(myf), mz = (int)floor(mzf);
  const int x1 = (int)floor(exf), y1 = (int)floor(eyf), z1 = (int)floor(ezf);
  // exact `(int)(x + offset_)` at the three corner voxels of the two half-rays
  const int pix = to_map_index(x0, g.offx), piy = to_map_index(y0, g.offy), piz = to_map_index(z0, g.offz);
  const int mix = to_map_index(mx, g.offx), miy = to_map_index(my, g.offy), miz = to_map_index(mz, g.offz);
  const int nix = to_map_index(x1, g.offx), niy = to_map_index(y1, g.offy), niz = to_map_index(z1, g.offz);
  auto inside = [&](int ix, int iy, int iz) { return (unsigned)ix < (unsigned)g.nx && (unsigned)iy < (unsigned)g.ny && (unsigned)iz < (unsigned)g.nz; };
  // fast path <=> the map-index conversion is a pure shift between the corners (it is monotone, so
  // equal deltas at the ends mean unit increments in between) and all corners are inside the map
  // (the marcher never leaves the box spanned by its end voxels).
  const bool fast = (fabs(mxf) + fabs(myf)) + fabs(mzf) < 1.0e9 &&        // finite (an infinite or NaN end takes the literal path)
                    inside(pix, piy, piz) && inside(mix, miy, miz) && inside(nix, niy, niz) &&
                    (mix - pix == mx - x0) && (miy - piy == my - y0) && (miz - piz == mz - z0) &&
                    (mix - nix == mx - x1) && (miy - niy == my - y1) && (miz - niz == mz - z1) &&
                    abs(mx - x0) + abs(my - y0) + abs(mz - z0) <= g.guard_steps && abs(mx - x1) + abs(my - y1) + abs(mz - z1) <= g.guard_steps;
  if (!fast) {
    slow_verdict = birc(g, vx, vy, vz, bx, by, bz, (flags & CONT_DISCARD) != 0, st) && (flags & CONT_INRANGE);
    return false;
  }
  c.vp = v; c.pt = pti; c.flags = flags;
  {
    const int dx = mx - x0, dy = my - y0, dz = mz - z0;
    // Certificate for the viewpoint-side half-ray.  A marcher whose start is strictly inside its voxel on every moving axis
    // makes exactly |dx| + |dy| + |dz| steps and never leaves the box spanned by its start and midpoint voxels (all its
    // boundary crossings have t < 1 by a margin no rounding can bridge; the reference's runaway cases start ON a lattice
    // plane).  If the summed-area table says that box holds no occupied voxel, every occCheck of this
    // side returns true whatever the exact voxel sequence is, so the side is not marched at all: its biNextId calls and
    // occCheck look-ups are only counted.
    bool pfree = false;
    if (g.sat != nullptr) {
      // "strictly inside its voxel on every moving axis": idle axes do not matter
      v_inside |= (dx == 0 ? 1 : 0) | (dy == 0 ? 2 : 0) | (dz == 0 ? 4 : 0);
      n_inside |= (mx == x1 ? 1 : 0) | (my == y1 ? 2 : 0) | (mz == z1 ? 4 : 0);
      const bool p_in = v_inside == 7, n_in = n_inside == 7;
      if (p_in && n_in) c.flags |= CONT_SANE;
      if (p_in)
        pfree = sat_box_count(g, min(pix, mix) >> FCVM_SAT_SHIFT, min(piy, miy) >> FCVM_SAT_SHIFT, min(piz, miz) >> FCVM_SAT_SHIFT, max(pix, mix) >> FCVM_SAT_SHIFT,
                              max(piy, miy) >> FCVM_SAT_SHIFT, max(piz, miz) >> FCVM_SAT_SHIFT) == 0;
    }
    c.p.rem = abs(dx) + abs(dy) + abs(dz);
    if (pfree) {
      c.flags |= CONT_PFREE;
      c.p.ix = mix; c.p.iy = miy; c.p.iz = miz;
      c.p.sx = c.p.sy = c.p.sz = 0;
      c.p.tmx = c.p.tmy = c.p.tmz = 0.0; c.p.tdx = c.p.tdy = c.p.tdz = 0.0;
    } else {
      c.p.ix = pix; c.p.iy = piy; c.p.iz = piz;
      c.p.sx = d_signum(dx); c.p.sy = d_signum(dy); c.p.sz = d_signum(dz);
      c.p.tmx = d_intbound_i(sxf, dx, g_rcp); c.p.tmy = d_intbound_i(syf, dy, g_rcp); c.p.tmz = d_intbound_i(szf, dz, g_rcp);
      c.p.tdx = t_delta(dx); c.p.tdy = t_delta(dy); c.p.tdz = t_delta(dz);
    }
  }
  {
    const int dx = mx - x1, dy = my - y1, dz = mz - z1;
    c.n.ix = nix; c.n.iy = niy; c.n.iz = niz;
    c.n.rem = abs(dx) + abs(dy) + abs(dz);
    c.n.sx = d_signum(dx); c.n.sy = d_signum(dy); c.n.sz = d_signum(dz);
    c.n.tmx = d_intbound_i(exf, dx, g_rcp); c.n.tmy = d_intbound_i(eyf, dy, g_rcp); c.n.tmz = d_intbound_i(ezf, dz, g_rcp);
    c.n.tdx = t_delta(dx); c.n.tdy = t_delta(dy); c.n.tdz = t_delta(dz);
  }
  if (flags & CONT_DISCARD) {
    // `raycaster_->biNextId(idx, idx_rev);` before the loop (cpp:387, 483, 845): both sides advance
    // one voxel (if they can), the reported voxels are not looked at
    if (c.flags & CONT_PFREE) c.p.rem -= c.p.rem > 0 ? 1 : 0;
    else c.p.step_if(c.p.rem > 0);
    c.n.step_if(c.n.rem > 0);
    c.flags &= ~CONT_DISCARD;
  }
  if (c.flags & CONT_PFREE) c.p.rem = max(0, c.p.rem - c.n.rem);      // calls left on the certified side once the target side has arrived
  return true;
}

// Conservative whole-box cull, shared by the tile level (a lane = a viewpoint) and the group level (a lane = 32 consecutive
// points): true when NO point of the axis-aligned box [lo, hi] can pass insideFOV for a viewpoint at (px, py, pz).  Either the
// squared distance from the viewpoint to the box exceeds the range band (every point fails `dir.norm() > max_dist`), or the
// whole box lies behind one of the four FOV planes: the largest d.n over the box is negative by a margin far above rounding,
// so every point fails that plane's `dir.dot(n) >= 0`.  Skipping such a box is exact.  nrm(q, k) = component k of normal q.
template <bool PLANES, class NormalAt>
__device__ __forceinline__ bool box_outside(const FovConst& fc, double px, double py, double pz, double lo0, double lo1, double lo2, double hi0,
                                            double hi1, double hi2, NormalAt nrm) {
  const double lx = lo0 - px, ly = lo1 - py, lz = lo2 - pz;
  const double hx = hi0 - px, hy = hi1 - py, hz = hi2 - pz;
  const double cx = fmax(fmax(lx, -hx), 0.0), cy = fmax(fmax(ly, -hy), 0.0), cz = fmax(fmax(lz, -hz), 0.0);
  bool culled = (cx * cx + cy * cy) + cz * cz > fc.md2_hi * 1.000001;
  if (PLANES && !culled) {
    const double margin = 1e-9 * ((fmax(fabs(lx), fabs(hx)) + fmax(fabs(ly), fabs(hy))) + fmax(fabs(lz), fabs(hz)) + 1.0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double n0 = nrm(q, 0), n1 = nrm(q, 1), n2 = nrm(q, 2);
      const double best = (fmax(lx * n0, hx * n0) + fmax(ly * n1, hy * n1)) + fmax(lz * n2, hz * n2);
      culled = culled || best < -margin;
    }
  }
  return culled;
}

template <int STAGE>
__global__ void __launch_bounds__(kWarps * 32, EVAL_MIN_CTAS) k_eval(const __grid_constant__ EvalArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  // [tile points][ready doubles][ready ints][entry queues]
  float4* tile = reinterpret_cast<float4*>(smem);
  double* ready_d = reinterpret_cast<double*>(smem + (size_t)a.tile_pts * sizeof(float4));
  int* ready_i = reinterpret_cast<int*>(ready_d + (size_t)kWarps * kContD * kReadyQ);
  QueueEntry* entries = reinterpret_cast<QueueEntry*>(ready_i + (size_t)kWarps * kContI * kReadyQ);
  __shared__ __align__(8) uint64_t bar;
  __shared__ int next_vp;
  __shared__ float4 gbox[2 * (EVAL_MAX_TILE / 32)];   // min / max corner of every group of 32 consecutive points of the tile
  __shared__ double aabb[6];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const long long tile_base = (long long)blockIdx.x * a.tile_pts;
  const int npts = (int)min((long long)a.tile_pts, a.P - tile_base);
  const int ngroups = (npts + 31) >> 5;                // <= 32: one lane per group at the group-level cull
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

  // ---- bounding boxes: one per group of 32 consecutive points, and the tile's (for the two conservative culls).  A group
  // that holds a non-finite coordinate gets an all-embracing box: such points are never culled (NaN compares false in the
  // reference's range test, so it goes on to cast the ray).
  {
    for (int gi = warp; gi < ngroups; gi += kWarps) {
      const int i = gi * 32 + lane;
      float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
      bool bad = false;
      if (i < npts) {
        const float4 q = tile[i];
        mn[0] = mx[0] = q.x; mn[1] = mx[1] = q.y; mn[2] = mx[2] = q.z;
        bad = !(fabsf(q.x) <= 3.0e38f && fabsf(q.y) <= 3.0e38f && fabsf(q.z) <= 3.0e38f);
      }
      bad = __any_sync(0xffffffffu, bad);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
          mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
      if (lane == 0) {
        gbox[2 * gi] = bad ? make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0.f) : make_float4(mn[0], mn[1], mn[2], 0.f);
        gbox[2 * gi + 1] = bad ? make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f) : make_float4(mx[0], mx[1], mx[2], 0.f);
      }
    }
    __syncthreads();
    if (warp == 0) {
      float4 lo = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f), hi = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0.f);
      if (lane < ngroups) { lo = gbox[2 * lane]; hi = gbox[2 * lane + 1]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
      }
      if (lane == 0) {
        aabb[0] = (double)lo.x; aabb[1] = (double)lo.y; aabb[2] = (double)lo.z;
        aabb[3] = (double)hi.x; aabb[4] = (double)hi.y; aabb[5] = (double)hi.z;
      }
    }
    __syncthreads();
  }

  const GridDesc& g = a.grid;
  const FovConst& fc = a.fov;
  QueueEntry* eq = entries + warp * kEntryQ;
  double* rqd = ready_d + (size_t)warp * kContD * kReadyQ;
  int* rqi = ready_i + (size_t)warp * kContI * kReadyQ;
  int en = 0;                       // warp-uniform entry-queue length
  int rn = 0;                       // warp-uniform ready-queue length
  RayStats st = {0u, 0u, 0u};
  unsigned int nrays = 0;
  unsigned long long npairs = 0;
  const long long W = a.words;

  auto set_bit = [&](uint32_t vp, uint32_t pt) {
    const long long gp = tile_base + pt;
    if (FCVM_CHECK(vp < (uint32_t)a.V && gp < a.P && (gp >> 5) < W)) atomicOr(a.rows + (long long)vp * W + (gp >> 5), 1u << (gp & 31));
  };

  // Pop `cnt` (<= 32) entries, run biInput for each on its own lane at full warp width, push the
  // rays that qualify for the fast path onto the ready queue; the rest are finished on the spot.
  auto prepare = [&](int cnt, Cont& c) -> bool {
    bool queued = false;
    if (lane < cnt) {
      const QueueEntry qe = eq[en - cnt + lane];
      nrays++;
      bool verdict = false;
#if defined(EVAL_EXP) && EVAL_EXP >= 2
      if (qe.pt == 0xFFFFFFFFu)   // timing experiment: FOV phase only
#else
      if (FCVM_CHECK(en - cnt + lane >= 0 && en - cnt + lane < kEntryQ && qe.pt < (uint32_t)npts && qe.vp < (uint32_t)a.V))
#endif
      {
        queued = prepare_ray<STAGE>(g, fc, a.recs, a.vp_ends, a.pt_ends, tile_base + qe.pt, qe.vp, qe.pt, tile[qe.pt], c, verdict, st);
        if (!queued && verdict) set_bit(qe.vp, qe.pt);
      }
    }
    en -= cnt;
#if defined(EVAL_EXP) && EVAL_EXP >= 1
    if (queued && c.p.rem == 123456789) set_bit(c.vp, c.pt);   // timing experiment: rays dropped after prepare
    queued = false;
#endif
    return queued;
  };
  // prepared rays onto the ready queue
  auto enqueue = [&](bool queued, const Cont& c) {
    const unsigned m = __ballot_sync(0xffffffffu, queued);
    if (queued && FCVM_CHECK(rn + __popc(m & lt_mask) < kReadyQ)) cont_store(rqd, rqi, rn + __popc(m & lt_mask), c);
    rn += __popc(m);
    __syncwarp();
  };

  // March rays from the ready queue, one per lane, both half-rays in lock step exactly as the call
  // sites drive biNextId.  A lane whose ray ends takes the next continuation at
  // once, so the warp stays full while rays of very different lengths finish.  Unless draining,
  // the warp suspends (writes the rays in flight back as continuations) as soon as the queue is
  // empty and fewer than `kKeep` lanes are busy, and goes back to testing FOV pairs.
  // EVAL_HANDOFF: the rays a warp has just prepared enter the march in the registers of their lanes (`active` / `c` on entry)
  // instead of through the queue; the idle lanes take queued (suspended) rays as before.
  auto march = [&](bool drain, bool active, Cont& c) {
    constexpr int kKeep = EVAL_KEEP;
    unsigned tick = 0;
    int pkey = -1, nkey = -1;
    unsigned long long pword = 0ull, nword = 0ull;
#if EVAL_STATS == 3
    int age = 0;                                      // measurement build: biNextId calls since the ray was (re)loaded
#endif
    for (;;) {
      const unsigned idle = __ballot_sync(0xffffffffu, !active);
      if (idle != 0u && rn > 0) {
        const int r = __popc(idle & lt_mask);
        if (!active && r < rn && FCVM_CHECK(rn - 1 - r >= 0 && rn - 1 - r < kReadyQ)) {
          cont_load(rqd, rqi, rn - 1 - r, c);
          active = true;
          pkey = -1; nkey = -1;
#if EVAL_STATS == 3
          age = 0;
#endif
        }
        rn -= min(__popc(idle), rn);
        __syncwarp();
      }
      const unsigned busy = __ballot_sync(0xffffffffu, active);
      if (busy == 0u) break;
      if (!drain && rn == 0 && __popc(busy) < kKeep) {
        if (active) cont_store(rqd, rqi, __popc(busy & lt_mask), c);
        rn = __popc(busy);
        __syncwarp();
        break;
      }
      // EVAL_UNROLL biNextId calls per refill / suspend check: the two ballots and the queue logic are
      // warp-wide overhead, a lane that finishes inside the group idles for at most EVAL_UNROLL - 1 steps
      // Every EVAL_QUERY_PERIOD-th round every ray in flight asks the summed-area table whether the box between its two fronts
      // (a certified viewpoint side's front is the midpoint voxel) still holds an occupied voxel.  If not, every look-up the
      // reference would still make returns "free": the pair is visible, its remaining biNextId calls are only counted.
      if (g.sat != nullptr && ((++tick) & (EVAL_QUERY_PERIOD - 1)) == 0) {
        if (active && (c.flags & CONT_SANE)) {
          const int n_empty = sat_box_count(g, min(c.p.ix, c.n.ix) >> FCVM_SAT_SHIFT, min(c.p.iy, c.n.iy) >> FCVM_SAT_SHIFT, min(c.p.iz, c.n.iz) >> FCVM_SAT_SHIFT,
                                            max(c.p.ix, c.n.ix) >> FCVM_SAT_SHIFT, max(c.p.iy, c.n.iy) >> FCVM_SAT_SHIFT, max(c.p.iz, c.n.iz) >> FCVM_SAT_SHIFT);
          if (n_empty == 0) {
            const int calls = (c.flags & CONT_PFREE) ? c.n.rem + c.p.rem : max(c.p.rem, c.n.rem);
            st.lookups += 2u * (unsigned)calls + 1u;
            if (c.flags & CONT_INRANGE) set_bit(c.vp, c.pt);
            active = false;
#if EVAL_STATS == 1
            st.trips++;                               // measurement build: rays ended by the in-flight certificate
#elif EVAL_STATS == 2
            st.walk += (unsigned)calls;               // measurement build: biNextId calls the certificate saved
#endif
          }
        }
      }
      const bool any_full = __any_sync(0xffffffffu, active && !(c.flags & CONT_PFREE));
#pragma unroll
      for (int u = 0; u < EVAL_UNROLL; ++u) {
        if (active) {
          // ---- one biNextId call and the loop body that follows it.  The call reports the voxels
          // BEFORE it steps, so they are looked up first and both sides advance afterwards.
          // A certified viewpoint side (CONT_PFREE) is never marched: p.ix/iy/iz stay at the midpoint voxel and p.rem holds
          // the number of biNextId calls that side would still make once the target side has arrived (each looks at two
          // free voxels).  In a warp that holds only such rays (any_full == false, the common case) the viewpoint side's
          // look-up and step are skipped altogether; in a mixed warp they run for every lane, harmlessly for the certified
          // ones (their side reports the free midpoint voxel and does not move).
          const bool pfree = (c.flags & CONT_PFREE) != 0;
          const bool pflag = !pfree && c.p.rem > 0, nflag = c.n.rem > 0;
          if (!(pflag || nflag)) {
            // the call returned false: visible_flag = occCheck(idx); visible_flag_rev keeps its last (true) value
            const bool f = occ_fast(g, c.p.ix, c.p.iy, c.p.iz, pkey, pword);
            st.lookups += pfree ? 1u + 2u * (unsigned)c.p.rem : 1u;
            // Both sides must have met in the midpoint voxel.  If they have not, the reference's marcher never reports
            // "at end": it walks on until an occCheck fails (an occupied voxel, or the map's edge), so the pair is NOT
            // visible there; counted as a cap trip (0 on every scene and test so far).
            const bool met = c.p.ix == c.n.ix && c.p.iy == c.n.iy && c.p.iz == c.n.iz;
            if (!met) st.trips++;
            if (met && f && (c.flags & CONT_INRANGE)) set_bit(c.vp, c.pt);
            active = false;
          } else {
            bool f = true;
            if (any_full) {
              f = occ_fast(g, c.p.ix, c.p.iy, c.p.iz, pkey, pword);
              c.p.step_if(pflag);
            }
            const bool r = occ_fast(g, c.n.ix, c.n.iy, c.n.iz, nkey, nword);
            st.lookups += 2;
            c.n.step_if(nflag);
#if EVAL_STATS == 1
            st.walk++;                                // measurement build: lane-steps marched
#elif EVAL_STATS == 3
            age++;
#endif
            if (!f || !r) {
              st.lookups += 1;                        // the re-check after the break (same voxel, same answer)
              active = false;
#if EVAL_STATS == 2
              st.trips++;                             // measurement build: rays ended by an occupied voxel
#elif EVAL_STATS == 3
              if (age == 1) st.walk++;                // measurement build: ... at the first call after loading / within the first three
              if (age <= 3) st.trips++;
#endif
            }
          }
        }
      }
    }
  };

  // Viewpoints are taken from the CTA's counter `chunk` at a time and culled against the tile's box with one viewpoint per
  // lane; the survivors are then culled against the 32-point groups' boxes with one group per lane, and only the groups
  // that remain are tested point by point.  Both culls are exact (box_outside), so the pair tests they skip are counted.
  const int chunk = max(1, min(32, a.vp_per_cta / (kWarps * 2)));
  for (;;) {
    int vi = 0;
    if (lane == 0) vi = atomicAdd(&next_vp, chunk);
    vi = __shfl_sync(0xffffffffu, vi, 0);
    const long long vbase = v_lo + vi;
    if (vbase >= v_hi) break;
    unsigned vmask;
    {
      const long long vl = vbase + lane;
      const bool mine = lane < chunk && vl < v_hi;
      bool culled = true;
      if (mine) {
        const VpRec* r = a.recs + vl;
        culled = box_outside<STAGE != 2>(fc, __ldg(&r->px), __ldg(&r->py), __ldg(&r->pz), aabb[0], aabb[1], aabb[2], aabb[3], aabb[4], aabb[5],
                                         [&](int q, int k) { return __ldg(&r->n[q][k]); });
        if (STAGE != 2 && culled) npairs += (unsigned long long)npts;
      }
      vmask = __ballot_sync(0xffffffffu, mine && !culled);
    }
   while (vmask != 0u) {
    const long long v = vbase + (__ffs(vmask) - 1);
    vmask &= vmask - 1u;

    // viewpoint record -> registers (uniform loads)
    VpRec rec;
    auto load_rec = [&]() {
      const double2* src = reinterpret_cast<const double2*>(a.recs + v);
      double2* dst = reinterpret_cast<double2*>(&rec);
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[k] = __ldg(src + k);
    };
    load_rec();
    const uint32_t* rrow = (STAGE == 2) ? a.restrict_rows + v * W + (tile_base >> 5) : nullptr;
    // groups worth a point-by-point test: E2 re-tests only what E1 found visible (a lane fetches its group's word of the E1
    // row: one coalesced load instead of 32 uniform ones); the others drop the groups whose box is out of range or behind a plane
    unsigned gmask;
    uint32_t rw_mine = 0u;
    if (STAGE == 2) {
      if (lane < ngroups) rw_mine = __ldg(rrow + lane);
      gmask = __ballot_sync(0xffffffffu, rw_mine != 0u);
    } else {
      bool keep = false;
      if (lane < ngroups) {
        const float4 lo = gbox[2 * lane], hi = gbox[2 * lane + 1];
        keep = !box_outside<true>(fc, rec.px, rec.py, rec.pz, (double)lo.x, (double)lo.y, (double)lo.z, (double)hi.x, (double)hi.y, (double)hi.z,
                                  [&](int q, int k) { return rec.n[q][k]; });
        if (!keep) npairs += (unsigned long long)min(32, npts - 32 * lane);
      }
      gmask = __ballot_sync(0xffffffffu, keep);
    }
    while (gmask != 0u) {
      const int grp = __ffs(gmask) - 1;
      gmask &= gmask - 1u;
      const int cb = grp << 5;
      const int i = cb + lane;
      bool valid = i < npts;
      if (STAGE == 2) {
        const uint32_t rw = __shfl_sync(0xffffffffu, rw_mine, grp);   // only points E1 found visible are re-tested
        valid = valid && ((rw >> lane) & 1u);
      }
      bool pass = false;
#if EVAL_STAGED_FOV
      // inside_fov (fcvm_device.cuh) unrolled into stages with warp-wide early outs: most groups of 32 consecutive
      // points are entirely out of range or entirely behind one FOV plane, and then the remaining dot products are
      // never computed.  A lane is dropped only where inside_fov would return false on either of its paths
      // (n2 > md2_hi; s[q] < -plane_eps), so the verdicts are unchanged.
      {
        double dx = 0.0, dy = 0.0, dz = 0.0, n2 = 0.0;
        bool alive = valid;
        if (valid) {
          const float4 pt = tile[i];
          dx = (double)pt.x - rec.px; dy = (double)pt.y - rec.py; dz = (double)pt.z - rec.pz;
          n2 = (dx * dx + dy * dy) + dz * dz;
          alive = !(n2 > fc.md2_hi);
          npairs++;
        }
        if (!__any_sync(0xffffffffu, alive)) continue;
        bool exact = !(n2 < fc.md2_lo);
        bool dead_group = false;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double sq = (dx * rec.n[q][0] + dy * rec.n[q][1]) + dz * rec.n[q][2];
          exact |= fabs(sq) <= fc.plane_eps;
          alive = alive && !(sq < -fc.plane_eps);
          if (!__any_sync(0xffffffffu, alive)) { dead_group = true; break; }
        }
        if (dead_group) continue;
        if (alive) {
          if (!exact) {
            pass = true;                      // all four s[q] > plane_eps and the range is clear of its band
          } else {
            // literal path (perception_utils.cpp:115-124)
            const double nrm = sqrt(n2);
            pass = !(nrm > fc.max_dist);
            if (pass) {
              if (n2 > 0.0) { dx = dx / nrm; dy = dy / nrm; dz = dz / nrm; }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const double d = (dx * rec.n[q][0] + dy * rec.n[q][1]) + dz * rec.n[q][2];
                if (d < 0.0) pass = false;
              }
            }
          }
        }
      }
#else
      if (valid) {
        const float4 pt = tile[i];
        pass = inside_fov(rec, fc, (double)pt.x, (double)pt.y, (double)pt.z);
        npairs++;
      }
#endif
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (m == 0u) continue;
      if (pass) {
        const int pos = en + __popc(m & lt_mask);
        if (FCVM_CHECK(pos < kEntryQ)) {
          eq[pos].vp = (uint32_t)v;
          eq[pos].pt = (uint32_t)i;
        }
      }
      en += __popc(m);
      __syncwarp();
      if (en >= 32) {
        Cont c;
        const bool queued = prepare(32, c);
#if EVAL_HANDOFF
        if (rn + __popc(__ballot_sync(0xffffffffu, queued)) >= 32) march(false, queued, c);
        else enqueue(queued, c);
#else
        enqueue(queued, c);
        if (rn >= 32) march(false, false, c);
#endif
        load_rec();   // re-read instead of keeping 30 registers live across the ray phases
      }
    }
   }
  }
  {
    Cont c;
    bool queued = false;
    if (en > 0) queued = prepare(en, c);   // en < 32 here
#if EVAL_HANDOFF
    march(true, queued, c);
#else
    enqueue(queued, c);
    march(true, false, c);
#endif
  }

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
  // rows are consumed strictly in order, but their words are fetched eight at a time: the sweep is a chain of dependent
  // L2 round trips otherwise (one warp-uniform 4-byte load per row)
  constexpr int kAhead = 8;
  for (int t0 = 0; t0 < nsched; t0 += kAhead) {
    int4 sv[kAhead];
    uint32_t wv[kAhead];
#pragma unroll
    for (int u = 0; u < kAhead; ++u) {
      const int t = min(t0 + u, nsched - 1);
      sv[u] = __ldg(sched + t);          // x = row, y = vp id, z = contri
      wv[u] = __ldg(rows + (long long)sv[u].x * words + col);
    }
#pragma unroll
    for (int u = 0; u < kAhead; ++u) {
      if (t0 + u >= nsched) break;
      const int4 s = sv[u];
      const uint32_t w = wv[u];
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
  }
  if (inb) { cover_contrib[j] = contrib; contrib_id[j] = owner; }
}

// ------------------------------------------------------------------------------------------
// Ownership as a per-point arg-max, the form that shards over GPUs.
// The rule of viewpoint_manager.cpp:500-528 / 589-617 applied to a whole stage is order-dependent only through ties:
// after all rows of the stage, point j belongs to the seer with the largest contribute_voxel_num, the earliest in
// processing order among equals - provided it beats (strictly) what j had before the stage, and j was not locked
// (owner id < last_vps_num).  Intermediate take-overs cancel in vps_voxcount (+1 then -1).  So:
//   k_own_best   key[j] = max over the rows a GPU holds of (contri << 32 | ~position-in-order), 0 = no seer
//   (N GPUs: one all-reduce MAX over the P keys - 8 MB at P = 10^6 - instead of gathering V x P / 8 bytes of rows)
//   k_own_apply  decode the winner, apply the rule once per point, +1 / -1 the two viewpoints' counts.
// ------------------------------------------------------------------------------------------
// sched[t] = {row, key_hi = contri, key_lo = ~position} for the local rows with contri > 0, sorted by key DESCENDING: the
// first row that sees a point is its local winner, so a thread only has to remember which of its 32 points are taken.
// The thread IS a 32-point word column: consecutive lanes read consecutive words of a row (128-byte coalesced loads,
// every row word read exactly once), stores happen once per point.
__global__ void __launch_bounds__(256) k_own_best(const uint32_t* __restrict__ rows, long long words, const int4* __restrict__ sched, int nsched,
                                                  long long P, unsigned long long* __restrict__ best) {
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col * 32 >= P) return;
  uint32_t taken = 0u;
  constexpr int kAhead = 8;
  for (int t0 = 0; t0 < nsched; t0 += kAhead) {
    int4 sv[kAhead];
    uint32_t wv[kAhead];
#pragma unroll
    for (int u = 0; u < kAhead; ++u) {
      sv[u] = __ldg(sched + min(t0 + u, nsched - 1));
      wv[u] = __ldg(rows + (long long)sv[u].x * words + col);
    }
#pragma unroll
    for (int u = 0; u < kAhead; ++u) {
      if (t0 + u >= nsched) break;
      uint32_t w = wv[u] & ~taken;
      if (w == 0u) continue;
      taken |= w;
      const unsigned long long key = ((unsigned long long)(uint32_t)sv[u].y << 32) | (unsigned long long)(uint32_t)sv[u].z;
      while (w) {
        const int bit = __ffs(w) - 1;
        w &= w - 1;
        const long long j = col * 32 + bit;
        if (j < P) best[j] = key;
      }
    }
    if (taken == 0xffffffffu) break;
  }
}

__global__ void __launch_bounds__(256) k_own_apply(const unsigned long long* __restrict__ best, const int* __restrict__ id_at_pos, int last_vps_num,
                                                   long long P, int* __restrict__ cover_contrib, int* __restrict__ contrib_id, int* __restrict__ voxcount) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P) return;
  const unsigned long long key = best[j];
  if (key == 0ull) return;
  const int c = (int)(key >> 32);
  const int vp = __ldg(id_at_pos + (uint32_t)(~(uint32_t)key));
  const int owner = contrib_id[j];
  if (owner < 0) {
    atomicAdd(voxcount + vp, 1);
  } else if (c > cover_contrib[j] && owner >= last_vps_num) {
    atomicSub(voxcount + owner, 1);
    atomicAdd(voxcount + vp, 1);
  } else {
    return;
  }
  cover_contrib[j] = c;
  contrib_id[j] = vp;
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <int STAGE>
static cudaError_t launch_eval_t(const EvalArgs& a, dim3 grid, size_t smem, cudaStream_t s) {
  // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute: remember it per device ordinal
  static std::atomic<bool> attr_done[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !attr_done[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(k_eval<STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eval_smem_bytes(EVAL_MAX_TILE));
    if (e != cudaSuccess) return e;
    k_init_rcp<<<(kRcpN + 255) / 256, 256, 0, s>>>();      // the device's reciprocal table (same stream: ordered before the first k_eval)
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev].store(true, std::memory_order_release);
  }
  k_eval<STAGE><<<grid, kWarps * 32, smem, s>>>(a);
  return cudaGetLastError();
}

cudaError_t preload_eval_kernels() {
  cudaFuncAttributes fa;
  const void* fns[] = {(const void*)k_init_rcp, (const void*)k_ray_ends, (const void*)k_eval<1>, (const void*)k_eval<2>, (const void*)k_eval<3>,
                       (const void*)k_eval<4>, (const void*)k_popc_rows, (const void*)k_exscan, (const void*)k_compact_rows, (const void*)k_own_sweep,
                       (const void*)k_own_best, (const void*)k_own_apply};
  for (const void* f : fns) {
    const cudaError_t e = cudaFuncGetAttributes(&fa, f);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

size_t eval_smem_bytes(int tile_pts) {
  return (size_t)tile_pts * sizeof(float4) + (size_t)kWarps * kReadyQ * (kContD * sizeof(double) + kContI * sizeof(int)) +
         (size_t)kWarps * kEntryQ * sizeof(QueueEntry);
}

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

cudaError_t launch_ray_ends(const double* xyz, int stride_d, const float4* pts, long long n, double res, RayEnd* out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_ray_ends<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(xyz, stride_d, pts, n, res, out);
  return cudaGetLastError();
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

cudaError_t launch_own_best(const uint32_t* rows, long long words, const int4* sched, int nsched, long long P, unsigned long long* best, cudaStream_t s) {
  if (nsched <= 0 || P <= 0) return cudaSuccess;
  const long long cols = (P + 31) / 32;
  k_own_best<<<(unsigned)((cols + 255) / 256), 256, 0, s>>>(rows, words, sched, nsched, P, best);
  return cudaGetLastError();
}

cudaError_t launch_own_apply(const unsigned long long* best, const int* id_at_pos, int last_vps_num, long long P, int* cover_contrib, int* contrib_id,
                             int* voxcount, cudaStream_t s) {
  if (P <= 0) return cudaSuccess;
  k_own_apply<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(best, id_at_pos, last_vps_num, P, cover_contrib, contrib_id, voxcount);
  return cudaGetLastError();
}

}  // namespace fcvm
