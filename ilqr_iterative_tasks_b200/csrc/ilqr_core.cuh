// ilqr_core.cuh -- device-side iLQR for one candidate terminal state (sm_100a, FP64).
//
// One lane owns one candidate.  Everything a candidate needs across iterations lives in a per-lane
// scratch column (shared memory for short horizons, an L2-resident global slab for long ones) laid out
// [element][lane] so that a warp's accesses to one element are contiguous; only the value function
// (Vx, Vxx) and the per-step working set are in registers.
//
// What is replaced (paths relative to the reference root):
//   ilqr            iterative_ilqr/control/iterative_ilqr.py:7-85
//   backward_pass   iterative_ilqr/control/iterative_ilqr.py:88-130
//   forward_pass    iterative_ilqr/control/iterative_ilqr.py:133-160
//   kinetic_bicycle / get_A_matrix / get_B_matrix   iterative_ilqr/systems/kinetic_bicycle.py:10-52
//   get_cost_derivation / add_control_constraint / repelling_cost_function / get_cost_final
//                   iterative_ilqr/control/ilqr_helper.py:9-64,83-150
//
// Structure that is exploited (the reference multiplies dense 4x4 / 4x2 arrays):
//   A_k = I + [[0,0,al,be],[0,0,ga,de],[0,0,0,0],[0,0,0,0]],  B_k = [[b0,0],[b1,0],[dt,0],[0,dt]],
//   Vxx symmetric (10 unique entries), l_uu diagonal and l_x / l_xx confined to the position block when
//   matrix_Q = matrix_R = 0 (the reference's default, kQR = false).
#pragma once
#ifndef I2LQR_HOST_BUILD
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdint.h>

#include "fastmath.cuh"

#ifdef I2LQR_HOST_BUILD  // developer screens compile the core for the host (tools/host_core_check.cpp): one "lane"
inline bool __any_sync(unsigned, bool p) { return p; }
#endif

namespace i2lqr {

// Parameter block in kernel-argument (constant-bank) space, derived on the host from i2lqr_params.
template <typename R>
struct KParamsT {
  int N;               // current horizon
  int N_param;         // ilqr_param.num_horizon (terminal obstacle advance, ilqr_helper.py:134-139)
  int K, L;            // num_ss_points, num_ss_iter
  int max_iter;        // max_ilqr_iter
  int max_relax;       // max_relax_iter
  int has_qr;          // matrix_Q != 0 or matrix_R != 0
  int pad0;
  R dt, dt2;      // timestep, timestep**2
  R hdt2;         // timestep**2 / 2: (u * dt2) / 2 and dt2 * c / 2 of the reference are u * hdt2 and hdt2 * c bit for bit
                  // (halving is exact and commutes with the product's rounding), one multiplication less each
  R a_max, d_lim; // clip bounds
  R a_min, d_min; // their negations (operands straight from the constant bank instead of a negation per use)
  R eps, lamb0, lamb_factor, max_lamb;
  R cg, ch, cq2;  // control barrier: q1*q2, q1*q2^2, q2
  R ca_lo, cd_lo; // exp(-2*q2*a_max), exp(-2*q2*delta_lim)
  R og, oh, oq2;  // obstacle barrier: q1*q2, q1*q2^2, q2
  R one_margin;   // 1 + safety_margin
  R xtarget[4];
  R Qf[10], Qf2[10];  // symmetric upper triangles of Qterminal and 2*Qterminal (00 01 02 03 11 12 13 22 23 33)
  R Q[10], Q2[10];    // matrix_Q and 2*matrix_Q
  R Rm[3], R2[3];     // matrix_R and 2*matrix_R (00 01 11)
};

using KParams = KParamsT<double>;

struct DevObstacle {  // mirrors i2lqr_obstacle (include/i2lqr.h)
  double x, y, width, height, spd;
  int32_t moving_option, present;
};

// Per-lane scratch column, element e at base[e * stride]: a "current trajectory" region (8 doubles per
// stage, ldc/stc) and a "gains / trial trajectory" region (10 doubles per stage, ldg/stg).
//   SmemScratch   both regions in shared memory; the stride is the block size, a compile-time constant, so
//                 element offsets fold into the LDS/STS immediates.
//   GlobalScratch both regions in an L2-resident global slab (long horizons).
//   HybridScratch current trajectory in shared memory, gains in the global slab: 384 instead of 864 bytes of
//                 shared memory per lane at N=6, i.e. 12 instead of 8 resident warps per SM.
// Scalar-addressed gains region shared by the memory-backed scratch types.
// Stage-loop protocol shared by all scratch types.  The backward pass, the forward pass and the accept copy walk
// the stages strictly in order; a scratch that STREAMS its slots through shared memory (StreamScratch) fetches
// ahead in begin / done and blocks in wait, for the others these are empty.  ldcs(k, i) is element i of the current
// trajectory's slot k as seen INSIDE those loops (a ring slot when streaming); ldc / stc address the backing store
// directly and are what the code outside the loops uses (candidate start, result, parking).
template <class Derived>
struct StageLoopDefaults {
  static constexpr bool kStream = false;
  __device__ __forceinline__ void bwd_begin(int) const {}
  __device__ __forceinline__ void bwd_wait(int) const {}
  __device__ __forceinline__ void bwd_done(int) const {}
  __device__ __forceinline__ void fwd_begin(int) const {}
  __device__ __forceinline__ void fwd_wait(int) const {}
  __device__ __forceinline__ void fwd_done(int) const {}
  __device__ __forceinline__ void cpy_begin(int) const {}
  __device__ __forceinline__ void cpy_wait(int) const {}
  __device__ __forceinline__ void cpy_done(int) const {}
  __device__ __forceinline__ auto ldcs(int k, int i) const { return static_cast<const Derived*>(this)->ldc(k * 8 + i); }
  // A candidate's cold state -- start state, cos / sin of its heading, terminal target: read once or twice per pass.
  // By default it lives in the candidate's registers; a placement that is short of registers (TmemLeanScratch) parks it.
  template <class C, typename R>
  __device__ __forceinline__ void load_x0(const C& c, R (&v)[4]) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = c.x0[i];
  }
  template <class C, typename R>
  __device__ __forceinline__ void load_cs0(const C& c, R& cs, R& sn) const {
    cs = c.c0;
    sn = c.s0;
  }
  template <class C, typename R>
  __device__ __forceinline__ void load_xt(const C& c, R (&v)[4]) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = c.xt[i];
  }
};
template <class Derived, typename R>
struct GainsByElement : StageLoopDefaults<Derived> {
  static constexpr bool kAligned = false;
  static constexpr bool kPrefetchCur = false;  // current-trajectory loads are cheap (shared memory)
  __device__ __forceinline__ void stg10(int k, const R (&v)[10]) const {
#pragma unroll
    for (int i = 0; i < 10; ++i) static_cast<const Derived*>(this)->stg(k * 10 + i, v[i]);
  }
  __device__ __forceinline__ void ldg10(int k, R (&v)[10]) const {
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = static_cast<const Derived*>(this)->ldg(k * 10 + i);
  }
  __device__ __forceinline__ void stg8(int k, const R (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) static_cast<const Derived*>(this)->stg(k * 10 + i, v[i]);
  }
  __device__ __forceinline__ void ldg8(int k, R (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = static_cast<const Derived*>(this)->ldg(k * 10 + i);
  }
  __device__ __forceinline__ void fence_gains() const {}
};
template <typename R, int kStride>
struct SmemScratch : GainsByElement<SmemScratch<R, kStride>, R> {
  using real = R;
  R* base;
  R* gbase;
  __device__ __forceinline__ void bind(R* smem_lane, R*, long long, int N) {
    base = smem_lane;
    gbase = smem_lane + (size_t)8 * N * kStride;
  }
  __device__ __forceinline__ R ldc(int e) const { return base[e * kStride]; }
  __device__ __forceinline__ void stc(int e, R v) const { base[e * kStride] = v; }
  __device__ __forceinline__ R ldg(int e) const { return gbase[e * kStride]; }
  __device__ __forceinline__ void stg(int e, R v) const { gbase[e * kStride] = v; }
  // gains / trial region of the lane `delta` positions away in the same warp (speculative lambda rungs)
  __device__ __forceinline__ R ldg_lane(int delta, int e) const { return gbase[e * kStride + delta]; }
};
template <typename R>
struct GlobalScratch : GainsByElement<GlobalScratch<R>, R> {
  using real = R;
  // the current trajectory lives in the L2 / HBM slab: the stage loops load the NEXT stage's operands while the
  // current stage computes, otherwise every stage starts with a dependent global-memory round trip
  static constexpr bool kPrefetchCur = true;
  R* base;
  R* gbase;
  long long stride;
  __device__ __forceinline__ void bind(R*, R* glob_lane, long long gstride, int N) {
    base = glob_lane;
    stride = gstride;
    gbase = glob_lane + (size_t)8 * N * gstride;
  }
  __device__ __forceinline__ R ldc(int e) const { return base[(size_t)e * stride]; }
  __device__ __forceinline__ void stc(int e, R v) const { base[(size_t)e * stride] = v; }
  __device__ __forceinline__ R ldg(int e) const { return gbase[(size_t)e * stride]; }
  __device__ __forceinline__ void stg(int e, R v) const { gbase[(size_t)e * stride] = v; }
  __device__ __forceinline__ R ldg_lane(int delta, int e) const { return gbase[(size_t)e * stride + delta]; }
};
template <typename R, int kStride>
struct HybridScratch : GainsByElement<HybridScratch<R, kStride>, R> {
  using real = R;
  R* base;
  R* gbase;
  long long stride;
  __device__ __forceinline__ void bind(R* smem_lane, R* glob_lane, long long gstride, int) {
    base = smem_lane;
    gbase = glob_lane;
    stride = gstride;
  }
  __device__ __forceinline__ R ldc(int e) const { return base[e * kStride]; }
  __device__ __forceinline__ void stc(int e, R v) const { base[e * kStride] = v; }
  __device__ __forceinline__ R ldg(int e) const { return gbase[(size_t)e * stride]; }
  __device__ __forceinline__ void stg(int e, R v) const { gbase[(size_t)e * stride] = v; }
  __device__ __forceinline__ R ldg_lane(int delta, int e) const { return gbase[(size_t)e * stride + delta]; }
};
#ifdef __CUDACC__
// TmemScratch: current trajectory in shared memory, gains / trial trajectory in Blackwell tensor memory
// (256 KB per SM next to the 228 KB of shared memory).  A warp owns one 32-lane quarter of TMEM; lane l of
// the warp addresses TMEM lane 32*(warp%4)+l, i.e. each thread gets a private strip of 32-bit columns:
// 20 columns (10 doubles) per stage.  tcgen05.ld / tcgen05.st are warp-collective (.sync.aligned), which is
// why ilqr_iteration is executed convergently by all 32 lanes.
template <int kStride>
struct TmemScratch : StageLoopDefaults<TmemScratch<kStride>> {
  using real = double;
  static constexpr bool kAligned = true;
  static constexpr bool kPrefetchCur = false;
  double* base;
  uint32_t taddr;  // (lane quarter << 16) | first column of this CTA's allocation
  __device__ __forceinline__ void bind_tmem(double* smem_lane, uint32_t tmem_addr) {
    base = smem_lane;
    taddr = tmem_addr;
  }
  __device__ __forceinline__ double ldc(int e) const { return base[e * kStride]; }
  __device__ __forceinline__ void stc(int e, double v) const { base[e * kStride] = v; }
  static __device__ __forceinline__ void st16(uint32_t a, const double* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :
        : "r"(a), "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
          "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3])),
          "r"(__double2loint(v[4])), "r"(__double2hiint(v[4])), "r"(__double2loint(v[5])), "r"(__double2hiint(v[5])),
          "r"(__double2loint(v[6])), "r"(__double2hiint(v[6])), "r"(__double2loint(v[7])), "r"(__double2hiint(v[7]))
        : "memory");
  }
  static __device__ __forceinline__ void st4(uint32_t a, const double* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 :
                 : "r"(a), "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1]))
                 : "memory");
  }
  static __device__ __forceinline__ void ld16(uint32_t a, double* v) {
    int r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __hiloint2double(r[2 * i + 1], r[2 * i]);
  }
  static __device__ __forceinline__ void ld4(uint32_t a, double* v) {
    int r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(a)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v[0] = __hiloint2double(r[1], r[0]);
    v[1] = __hiloint2double(r[3], r[2]);
  }
  __device__ __forceinline__ void stg10(int k, const double (&v)[10]) const {
    st16(taddr + k * 20, v);
    st4(taddr + k * 20 + 16, v + 8);
  }
  __device__ __forceinline__ void stg8(int k, const double (&v)[8]) const { st16(taddr + k * 20, v); }
  __device__ __forceinline__ void ldg10(int k, double (&v)[10]) const {
    ld16(taddr + k * 20, v);
    ld4(taddr + k * 20 + 16, v + 8);
  }
  __device__ __forceinline__ void ldg8(int k, double (&v)[8]) const { ld16(taddr + k * 20, v); }
  // a tcgen05.ld must not overtake the tcgen05.st that produced its data
  __device__ __forceinline__ void fence_gains() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
};
// TmemLeanScratch (horizons up to 6, flat batches): the tensor-memory placement with the candidate's cold state parked
// as well -- start state and cos / sin of its heading in six more doubles of shared memory per lane, the terminal target
// in the eight tensor-memory columns that the gains leave free of the CTA's 128 -- so that the kernel fits 128 registers
// and FOUR CTAs (16 warps) share an SM instead of three.  Same arithmetic, same operands, same bits.
template <int kStride>
struct TmemLeanScratch : TmemScratch<kStride> {
  using Base = TmemScratch<kStride>;
  double* cold;     // shared: [6][kStride] behind the trajectory slots
  uint32_t txt;     // tensor-memory address of the terminal target (4 doubles)
  __device__ __forceinline__ void bind_lean(double* smem_lane, uint32_t tmem_addr, int N) {
    Base::bind_tmem(smem_lane, tmem_addr);
    cold = smem_lane + (size_t)8 * N * kStride;
    txt = tmem_addr + 20 * N;
  }
  template <class C>
  __device__ __forceinline__ void park_start(const C& c) const {  // any subset of lanes
#pragma unroll
    for (int i = 0; i < 4; ++i) cold[i * kStride] = c.x0[i];
    cold[4 * kStride] = c.c0;
    cold[5 * kStride] = c.s0;
  }
  // all 32 lanes: lanes with `fresh` park the terminal target they hold in registers, the others keep theirs
  template <class C>
  __device__ __forceinline__ void park_target(const C& c, bool fresh) const {
    double v[4];
    ld_xt(v);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = fresh ? c.xt[i] : v[i];
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :
                 : "r"(txt), "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
                   "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3]))
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  __device__ __forceinline__ void ld_xt(double (&v)[4]) const {
    int r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(txt)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __hiloint2double(r[2 * i + 1], r[2 * i]);
  }
  template <class C>
  __device__ __forceinline__ void load_x0(const C&, double (&v)[4]) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = cold[i * kStride];
  }
  template <class C>
  __device__ __forceinline__ void load_cs0(const C&, double& cs, double& sn) const {
    cs = cold[4 * kStride];
    sn = cold[5 * kStride];
  }
  template <class C>
  __device__ __forceinline__ void load_xt(const C&, double (&v)[4]) const { ld_xt(v); }  // warp-collective
};
#endif

// StreamScratch (long horizons): a lane's working set lives in global memory, laid out per WARP as
//   CUR [N][6][32 lanes]   current trajectory slots        NEW [N][10][32 lanes]   gains / trial trajectory slots
// so that one stage of one region is a contiguous 1.5 KB / 2.5 KB block.  All three stage loops walk the stages in
// order, so the warp streams them through a shared-memory ring of kRing slots with bulk asynchronous copies
// (cp.async.bulk global -> shared, completion on an mbarrier per slot: the TMA engine moves the data, one lane
// issues): the loads of the next kRing - 2 stages are in flight while a stage computes, and the loop bodies read
// shared memory.  Results (gains, trial, accepted trajectory) go straight to global memory with coalesced stores,
// 256 B per element per warp.  The kernel is then bound by HBM bandwidth (measured at N = 50: 4.8 TB/s of DRAM
// traffic) instead of by dependent L2 / HBM round trips per stage -- so the slots carry no cos / sin of the heading
// (6 doubles: u_k, x_{k+1}): the backward pass recomputes them with the same routine from the same heading, bit for
// bit what the other placements store, and a sixth of the traffic is gone.
template <int kRing>
struct StreamScratch : StageLoopDefaults<StreamScratch<kRing>> {
  using real = double;
  static constexpr bool kAligned = true;  // mbarrier waits and bulk copies are warp-wide: every lane runs the iteration
  static constexpr bool kPrefetchCur = false;
  static constexpr bool kStream = true;
  static constexpr int kCur = 6;                  // doubles per current-trajectory / trial slot
  static constexpr int kSlot = (kCur + 10) * 32;  // doubles per ring slot: CUR part [6][32], then NEW part [10][32]
  static constexpr unsigned kCurBytes = kCur * 32 * 8, kNewBytes = 10 * 32 * 8;
  static_assert((kRing & (kRing - 1)) == 0 && kRing >= 4, "ring size: a power of two, at least 4");
  double* ring;    // shared
  uint32_t bars;   // shared address of kRing mbarriers
  double* gcur;    // global
  double* gnew;    // global
  int lane, N;
  mutable unsigned phase;  // expected parity per slot (the scratch is handed around as a const reference)
#ifdef __CUDACC__
  __device__ __forceinline__ void bind_stream(double* ring_smem, double* gwarp, int lane_, int N_) {
    ring = ring_smem;
    bars = (uint32_t)__cvta_generic_to_shared(ring_smem + (size_t)kRing * kSlot);
    gcur = gwarp;
    gnew = gwarp + (size_t)N_ * kCur * 32;
    lane = lane_;
    N = N_;
    phase = 0;
    if (lane == 0) {
      for (int i = 0; i < kRing; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * i) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  // ---- backing store, direct.  e = 8 k + i as everywhere else; i = 6, 7 (cos / sin of the heading) are not stored.
  __device__ __forceinline__ double ldc(int e) const {
    const int k = e >> 3, i = e & 7;
    if (i < kCur) return gcur[((size_t)k * kCur + i) * 32 + lane];
    double sn, cs;
    sincos_fast(gcur[((size_t)k * kCur + 5) * 32 + lane], &sn, &cs);
    return i == 6 ? cs : sn;
  }
  __device__ __forceinline__ void stc(int e, double v) const {
    const int k = e >> 3, i = e & 7;
    if (i < kCur) gcur[((size_t)k * kCur + i) * 32 + lane] = v;
  }
  __device__ __forceinline__ void stg10(int k, const double (&v)[10]) const {
#pragma unroll
    for (int i = 0; i < 10; ++i) gnew[((size_t)k * 10 + i) * 32 + lane] = v[i];
  }
  __device__ __forceinline__ void stg8(int k, const double (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < kCur; ++i) gnew[((size_t)k * 10 + i) * 32 + lane] = v[i];
  }
  __device__ __forceinline__ void fence_gains() const {}
  // The finished trajectory of lane `j` in the reference's layout, xvar (4, N+1) and uvar (2, N), written by the WHOLE warp
  // (lane l takes stages l, l + 32, ...): a lane copying its own 6 N values is a chain of 6 N dependent global loads, each
  // followed by a store, during which the other 31 candidates of the warp stand still (measured: a third of the kernel).
  __device__ __forceinline__ void store_lane(int j, const double (&x0)[4], double* __restrict__ xv, double* __restrict__ uv) const {
    for (int k0 = lane; k0 < N; k0 += 64) {  // two stages per lane and trip: twelve loads in flight before the first store
      const int k1 = k0 + 32;
      const bool two = k1 < N;
      const double* s0 = gcur + (size_t)k0 * kCur * 32 + j;
      const double* s1 = gcur + (size_t)(two ? k1 : k0) * kCur * 32 + j;
      double a[kCur], b[kCur];
#pragma unroll
      for (int i = 0; i < kCur; ++i) {
        a[i] = __ldcg(s0 + i * 32);
        b[i] = __ldcg(s1 + i * 32);
      }
      if (uv != nullptr) {
        uv[k0] = a[0];
        uv[N + k0] = a[1];
        if (two) {
          uv[k1] = b[0];
          uv[N + k1] = b[1];
        }
      }
      if (xv != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          xv[i * (N + 1) + k0 + 1] = a[2 + i];
          if (two) xv[i * (N + 1) + k1 + 1] = b[2 + i];
        }
      }
    }
    if (lane == 0 && xv != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i * (N + 1)] = x0[i];
    }
  }
  // ---- ring
  __device__ __forceinline__ double ldcs(int k, int i) const { return ring[(k & (kRing - 1)) * kSlot + i * 32 + lane]; }  // i < 6
  __device__ __forceinline__ void ldg10(int k, double (&v)[10]) const {
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = ring[(k & (kRing - 1)) * kSlot + (kCur + i) * 32 + lane];
  }
  __device__ __forceinline__ void ldg8(int k, double (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < kCur; ++i) v[i] = ring[(k & (kRing - 1)) * kSlot + (kCur + i) * 32 + lane];
    v[6] = v[7] = 0.0;
  }
  // what to fetch of a stage: 1 = CUR slot, 2 = NEW slot (10 elements), 4 = the trial part of the NEW slot (6 elements)
  __device__ __forceinline__ void issue(int k, int what) const {
    if (lane != 0) return;
    const int slot = k & (kRing - 1);
    const uint32_t bar = bars + 8 * slot;
    const uint32_t bytes = ((what & 1) ? kCurBytes : 0u) + ((what & 2) ? kNewBytes : 0u) + ((what & 4) ? kCurBytes : 0u);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + (size_t)slot * kSlot);
    if (what & 1)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(gcur + (size_t)k * kCur * 32), "r"(kCurBytes), "r"(bar)
                   : "memory");
    if (what & 6)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + kCurBytes),
                   "l"(gnew + (size_t)k * 10 * 32), "r"((what & 2) ? kNewBytes : kCurBytes), "r"(bar)
                   : "memory");
  }
  __device__ __forceinline__ void wait(int k) const {
    const int slot = k & (kRing - 1);
    const uint32_t bar = bars + 8 * slot, par = (phase >> slot) & 1u;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(par)
        : "memory");
    phase ^= 1u << slot;
  }
  // Everything the warp wrote to global memory with ordinary stores (candidate start, gains, trial, accepted slots)
  // must be visible to the bulk copies (async proxy) that read it back, and nobody may still be reading a ring slot
  // that the next pass refills.
  __device__ __forceinline__ void pass_fence() const {
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
  }
  // backward: stages N-1 ... 0, stage k reads CUR slots k and k-1
  __device__ __forceinline__ void bwd_begin(int) const {
    pass_fence();
    for (int k = N - 1; k >= 0 && k > N - 1 - kRing; --k) issue(k, 1);
    wait(N - 1);
  }
  __device__ __forceinline__ void bwd_wait(int k) const {
    if (k > 0) wait(k - 1);
  }
  __device__ __forceinline__ void bwd_done(int k) const {
    __syncwarp();
    if (k - kRing >= 0) issue(k - kRing, 1);
  }
  // forward: stages 0 ... N-1, stage k reads CUR slots k-1, k and NEW slots k, k+1
  __device__ __forceinline__ void fwd_begin(int) const {
    pass_fence();
    for (int k = 0; k < N && k < kRing; ++k) issue(k, 3);
    wait(0);
  }
  __device__ __forceinline__ void fwd_wait(int k) const {
    if (k + 1 < N) wait(k + 1);
  }
  __device__ __forceinline__ void fwd_done(int k) const {
    __syncwarp();
    if (k >= 1 && k - 1 + kRing < N) issue(k - 1 + kRing, 3);
  }
  // accept copy: stages 0 ... N-1, stage k reads the trial part of NEW slots k (already in registers) and k+1
  __device__ __forceinline__ void cpy_begin(int) const {
    pass_fence();
    for (int k = 0; k < N && k < kRing; ++k) issue(k, 4);
    wait(0);
  }
  __device__ __forceinline__ void cpy_wait(int k) const {
    if (k + 1 < N) wait(k + 1);
  }
  __device__ __forceinline__ void cpy_done(int k) const {
    __syncwarp();
    if (k + kRing < N) issue(k + kRing, 4);
  }
#endif
};
// Scratch layout for horizon N (18 N doubles per candidate):
//   cur slot k (8):  [0..1] u_k   [2..5] x_{k+1}   [6] cos(theta_{k+1})  [7] sin(theta_{k+1})
//   new slot k (10): gains [0..1] k_k, [2..9] K_k (row-major 2x4) after the backward pass; the forward pass
//                    consumes the gains of step k and overwrites the slot with the trial trajectory in the
//                    cur-slot format (first 8 entries).
__host__ __device__ __forceinline__ int scratch_doubles(int N) { return 18 * N; }
#define I2_CUR(k) ((k)*8)
#define I2_NEW(N, k) ((N)*8 + (k)*10)

// Registers that persist over a candidate's iterations.
template <typename R>
struct CandT {
  R x0[4], c0, s0;  // initial state, cos / sin of its heading
  R xt[4];          // terminal target (safe-set point)
  R lamb, J;        // LM regularisation, cost of the current trajectory
  // obstacle: centre, 1/width^2, 1/height^2, per-stage drift of the centre (left / up), barrier scales.
  // `obstacle is None` is encoded as zero scales and zero P so that the same straight-line code runs.
  R ox, oy, oP0, oP1, osx, osy, og, oh;
  int iters;
  // further obstacles of the episode (extension, controller-step kernels only: SURVEY.md 8 f2); their barrier terms add
  const DevObstacle* ex;
  int n_ex;
};
constexpr int kMaxExtraObstacles = 3;  // an episode's obstacle list holds up to 1 + 3 ellipses
using Cand = CandT<double>;

template <typename R>
__device__ __forceinline__ R clipd(R v, R lo, R hi) { return v < lo ? lo : (v > hi ? hi : v); }

// v^T M v for symmetric M stored as upper triangle
template <typename R>
__device__ __forceinline__ R quad4(const R* M, R a, R b, R c, R d) {
  R r0 = M[0] * a + M[1] * b + M[2] * c + M[3] * d;
  R r1 = M[1] * a + M[4] * b + M[5] * c + M[6] * d;
  R r2 = M[2] * a + M[5] * b + M[7] * c + M[8] * d;
  R r3 = M[3] * a + M[6] * b + M[8] * c + M[9] * d;
  return r0 * a + r1 * b + r2 * c + r3 * d;
}
template <typename R>
__device__ __forceinline__ R quad2(const R* M, R a, R b) {
  return (M[0] * a + M[1] * b) * a + (M[1] * a + M[2] * b) * b;
}

// Obstacle ellipse barrier at position (px,py), `step` stages ahead: gradient (2) and Gauss-Newton Hessian
// (00,01,11) of q1*exp(q2*h), h = 1 + margin - d^T P d   (ilqr_helper.py:32-51,118-149, :59-64).
// The centre is (x_o - step*spd, y_o) when moving left, (x_o, y_o + step*spd) when moving up.
template <typename R>
__device__ __forceinline__ void obstacle_barrier(const KParamsT<R>& p, const CandT<R>& c, R px, R py, int step,
                                                 R& g0, R& g1, R& h00, R& h01, R& h11) {
  const R st = R(step);
  const R cx = c.ox - st * c.osx, cy = c.oy + st * c.osy;
  const R d0 = px - cx, d1 = py - cy;
  const R h = p.one_margin - ((d0 * c.oP0) * d0 + (d1 * c.oP1) * d1);
  const R hd0 = (-R(2.0) * c.oP0) * d0, hd1 = (-R(2.0) * c.oP1) * d1;
  const R e = exp_fast_lo(p.oq2 * h);  // h <= 1 + margin: only underflow needs guarding
  const R c1 = c.og * e, c2 = c.oh * e;
  g0 = c1 * hd0;
  g1 = c1 * hd1;
  h00 = c2 * (hd0 * hd0);
  h01 = c2 * (hd0 * hd1);
  h11 = c2 * (hd1 * hd1);
}

// The same barrier for the further obstacles of an episode's list, ADDED to the primary's terms: the reference's cost
// loop takes one Obstacle (ilqr_helper.py:16-24,32-51); a list is priced as the sum of its members' barriers, each
// evaluated exactly as the single one (same association, 1 / w^2 by division as load_obstacle does).
template <typename R>
__device__ __forceinline__ void obstacle_barrier_extra(const KParamsT<R>& p, const CandT<R>& c, R px, R py, int step, R& g0, R& g1, R& h00,
                                                       R& h01, R& h11) {
  for (int m = 0; m < c.n_ex; ++m) {
    const DevObstacle o = c.ex[m];
    if (!o.present) continue;
    const R st = R(step);
    const R osx = R(o.moving_option == 2 ? o.spd : 0.0), osy = R(o.moving_option == 1 ? o.spd : 0.0);
    const R P0 = R(1.0 / (o.width * o.width)), P1 = R(1.0 / (o.height * o.height));
    const R cx = R(o.x) - st * osx, cy = R(o.y) + st * osy;
    const R d0 = px - cx, d1 = py - cy;
    const R h = p.one_margin - ((d0 * P0) * d0 + (d1 * P1) * d1);
    const R hd0 = (-R(2.0) * P0) * d0, hd1 = (-R(2.0) * P1) * d1;
    const R e = exp_fast_lo(p.oq2 * h);
    const R c1 = p.og * e, c2 = p.oh * e;
    g0 += c1 * hd0;
    g1 += c1 * hd1;
    h00 += c2 * (hd0 * hd0);
    h01 += c2 * (hd0 * hd1);
    h11 += c2 * (hd1 * hd1);
  }
}

// Regularised inverse of the 2x2 Quu (iterative_ilqr.py:118-123): eigen-decompose, clamp negative
// eigenvalues to zero, add lamb, V diag(1/e) V^T.  When both eigenvalues are non-negative (always, in the
// reference's own runs: SURVEY.md 8a, a5) the clamp is the identity and the result is exactly
// (Quu + lamb I)^-1, evaluated here in closed form with one reciprocal.  Otherwise the symmetric
// eigen-decomposition is done by one Jacobi rotation (stable when |b| << |a-d|).
#ifdef __CUDACC__
#define I2_NOINLINE __device__ __noinline__
#else
#define I2_NOINLINE inline
#endif
// (the rare branch, kept out of line: inlined it is 200 instructions in the middle of every backward stage)
template <typename R>
struct Sym2 {
  R m00, m01, m11;
};
template <typename R>
I2_NOINLINE Sym2<R> quu_inverse_clamped(R a, R b, R d, R lamb) {
  R c = R(1.0), s = R(0.0), e1 = a, e2 = d;
  if (b != R(0.0)) {
    R tau = (d - a) / (R(2.0) * b);
    R t = (tau >= R(0.0) ? R(1.0) : -R(1.0)) / (fabs(tau) + sqrt(R(1.0) + tau * tau));
    c = R(1.0) / sqrt(R(1.0) + t * t);
    s = t * c;
    e1 = a - t * b;
    e2 = d + t * b;
  }
  e1 = (e1 < R(0.0) ? R(0.0) : e1) + lamb;
  e2 = (e2 < R(0.0) ? R(0.0) : e2) + lamb;
  R r1 = R(1.0) / e1, r2 = R(1.0) / e2;
  Sym2<R> o;
  o.m00 = c * r1 * c + s * r2 * s;
  o.m01 = s * r2 * c - c * r1 * s;
  o.m11 = s * r1 * s + c * r2 * c;
  return o;
}
template <typename R>
__device__ __forceinline__ void quu_inverse(R a, R b, R d, R lamb, R& i00, R& i01,
                                            R& i11) {
  if (a * d - b * b >= R(0.0) && a + d >= R(0.0)) {
    const R al = a + lamb, dl = d + lamb;
    const R r = rcp_fast(al * dl - b * b);
    i00 = dl * r;
    i01 = -b * r;
    i11 = al * r;
    return;
  }
  const Sym2<R> o = quu_inverse_clamped(a, b, d, lamb);
  i00 = o.m00;
  i01 = o.m01;
  i11 = o.m11;
}

// Terminal cost (x_N - x_t)^T Qf (x_N - x_t)
template <typename R>
__device__ __forceinline__ R terminal_cost(const KParamsT<R>& p, const CandT<R>& c, const R* xN) {
  return quad4(p.Qf, xN[0] - c.xt[0], xN[1] - c.xt[1], xN[2] - c.xt[2], xN[3] - c.xt[3]);
}

template <typename R>
__device__ __forceinline__ R terminal_cost_to(const KParamsT<R>& p, const R (&xt)[4], const R* xN) {
  return quad4(p.Qf, xN[0] - xt[0], xN[1] - xt[1], xN[2] - xt[2], xN[3] - xt[3]);
}

// Candidate start (utils/base.py:405-410 + the first rollout of iterative_ilqr.py:32-48).  With u0 == null the
// inputs are zero (what calc_input passes): heading and speed stay put and every stage reuses
// cos/sin(theta_0).  u0 (2,N) is the warm start of a direct `ilqr()` call.
template <bool kQR, int kN, class Scratch, typename R = typename Scratch::real>
__device__ __forceinline__ void init_candidate(const KParamsT<R>& p, CandT<R>& c, const Scratch& s, const double* u0) {
  const int N = kN > 0 ? kN : p.N;  // kN > 0: horizon known at compile time, stage loops fully unrolled
  R x[4] = {c.x0[0], c.x0[1], c.x0[2], c.x0[3]};
  R cs = c.c0, sn = c.s0;
  R J = R(0.0);
#pragma unroll(kN > 0 ? kN : 1)
  for (int k = 0; k < N; ++k) {
    R ua = R(0.0), ud = R(0.0);
    if (u0 != nullptr) {
      ua = clipd(R(u0[k]), p.a_min, p.a_max);
      ud = clipd(R(u0[N + k]), p.d_min, p.d_lim);
    }
    if (kQR) {
      const R l_state = quad4(p.Q, x[0] - p.xtarget[0], x[1] - p.xtarget[1], x[2] - p.xtarget[2], x[3] - p.xtarget[3]);
      J = J + l_state + quad2(p.Rm, ua, ud);
    }
    const R sd = x[2] * p.dt + ua * p.hdt2;
    x[0] = x[0] + cs * sd;
    x[1] = x[1] + sn * sd;
    x[2] = x[2] + ua * p.dt;
    x[3] = x[3] + ud * p.dt;
    if (u0 != nullptr) sincos_fast(x[3], &sn, &cs);
    s.stc(I2_CUR(k) + 0, ua);
    s.stc(I2_CUR(k) + 1, ud);
    s.stc(I2_CUR(k) + 2, x[0]);
    s.stc(I2_CUR(k) + 3, x[1]);
    s.stc(I2_CUR(k) + 4, x[2]);
    s.stc(I2_CUR(k) + 5, x[3]);
    s.stc(I2_CUR(k) + 6, cs);
    s.stc(I2_CUR(k) + 7, sn);
  }
  c.J = J + terminal_cost(p, c, x);
  c.iters = 0;
}

// Backward Riccati pass over the current trajectory; leaves the gains in the new-slots.
template <bool kQR, int kN, class Scratch, typename R = typename Scratch::real>
__device__ __forceinline__ void backward_pass(const KParamsT<R>& p, const CandT<R>& c, const Scratch& s) {
  const int N = kN > 0 ? kN : p.N;  // kN > 0: horizon known at compile time, stage loops fully unrolled
  const R dt = p.dt;
  // ---- value function at the horizon end: get_cost_final (ilqr_helper.py:106-150)
  s.bwd_begin(N);
  R xN[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) xN[i] = s.ldcs(N - 1, 2 + i);
  R Vx0, Vx1, Vx2, Vx3;
  {
    R xt[4];
    s.load_xt(c, xt);
    R e0 = xN[0] - xt[0], e1 = xN[1] - xt[1], e2 = xN[2] - xt[2], e3 = xN[3] - xt[3];
    Vx0 = p.Qf2[0] * e0 + p.Qf2[1] * e1 + p.Qf2[2] * e2 + p.Qf2[3] * e3;
    Vx1 = p.Qf2[1] * e0 + p.Qf2[4] * e1 + p.Qf2[5] * e2 + p.Qf2[6] * e3;
    Vx2 = p.Qf2[2] * e0 + p.Qf2[5] * e1 + p.Qf2[7] * e2 + p.Qf2[8] * e3;
    Vx3 = p.Qf2[3] * e0 + p.Qf2[6] * e1 + p.Qf2[8] * e2 + p.Qf2[9] * e3;
  }
  R V00 = p.Qf2[0], V01 = p.Qf2[1], V02 = p.Qf2[2], V03 = p.Qf2[3], V11 = p.Qf2[4], V12 = p.Qf2[5],
         V13 = p.Qf2[6], V22 = p.Qf2[7], V23 = p.Qf2[8], V33 = p.Qf2[9];
  {
    R g0, g1, h00, h01, h11;
    obstacle_barrier(p, c, xN[0], xN[1], p.N_param, g0, g1, h00, h01, h11);
    Vx0 += g0;
    Vx1 += g1;
    V00 += h00;
    V01 += h01;
    V11 += h11;
  }
  // operands of one stage: u_k, v_{k+1}, cos / sin theta_{k+1} from slot k, x_k from slot k-1 (x_0 for k = 0)
  auto load_stage = [&](int k, R (&o)[9]) {
    o[0] = s.ldcs(k, 0);
    o[1] = s.ldcs(k, 1);
    o[2] = s.ldcs(k, 4);
    if constexpr (Scratch::kStream) {  // (the streamed slots carry no cos / sin: same routine, same heading, same bits)
      sincos_fast(R(s.ldcs(k, 5)), &o[4], &o[3]);
    } else {
      o[3] = s.ldcs(k, 6);
      o[4] = s.ldcs(k, 7);
    }
    if (k == 0) {
      R x0[4];
      s.load_x0(c, x0);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[5 + i] = x0[i];
    } else {
      o[5] = s.ldcs(k - 1, 2);
      o[6] = s.ldcs(k - 1, 3);
      if (kQR) {
        o[7] = s.ldcs(k - 1, 4);
        o[8] = s.ldcs(k - 1, 5);
      } else {
        o[7] = o[8] = R(0.0);
      }
    }
  };
  R pf[9];
  if (Scratch::kPrefetchCur) load_stage(N - 1, pf);
#pragma unroll(kN > 0 ? kN : 1)  // (two stages per trip measured 11 % slower here: 720-instruction body)
  for (int k = N - 1; k >= 0; --k) {
    // ---- trajectory at stage k
    s.bwd_wait(k);
    R op[9];
    if (Scratch::kPrefetchCur) {
#pragma unroll
      for (int i = 0; i < 9; ++i) op[i] = pf[i];
      if (k > 0) load_stage(k - 1, pf);  // in flight while this stage computes
    } else {
      load_stage(k, op);
    }
    const R ua = op[0], ud = op[1];
    const R v1 = op[2];                 // v_{k+1}
    const R cs = op[3], sn = op[4];     // cos / sin theta_{k+1}
    R xk[4] = {op[5], op[6], op[7], op[8]};
    // ---- linearisation at (v_{k+1}, theta_{k+1}, a_k): iterative_ilqr.py:92-99, kinetic_bicycle.py:30-52
    const R sp = v1 * dt + ua * p.hdt2;
    const R al = cs * dt, be = -sp * sn, ga = sn * dt, de = sp * cs;
    const R b0 = p.hdt2 * cs, b1 = p.hdt2 * sn;
    // ---- stage cost derivatives: ilqr_helper.py:25-55, input barriers :83-103
    R lu0, lu1, luu00, luu11;
    {
      // the lower-bound barriers are the reciprocals of the upper-bound ones up to a constant:
      // exp(q2(-ubar - u)) = exp(-2 q2 ubar) / exp(q2(u - ubar)) -- two exp + two reciprocals instead of four exp;
      // u is clipped, so the exponents lie in [-2 q2 ubar, 0] and need no range guard
      const R e1 = exp_fast_bounded(p.cq2 * (ua - p.a_max)), e3 = exp_fast_bounded(p.cq2 * (ud - p.d_lim));
      const R e2 = p.ca_lo * rcp_fast(e1), e4 = p.cd_lo * rcp_fast(e3);
      lu0 = p.cg * e1 - p.cg * e2;
      lu1 = p.cg * e3 - p.cg * e4;
      luu00 = p.ch * e1 + p.ch * e2;
      luu11 = p.ch * e3 + p.ch * e4;
    }
    R g0, g1, l00, l01, l11;  // obstacle barrier: gradient and Hessian of the position block
    obstacle_barrier(p, c, xk[0], xk[1], k, g0, g1, l00, l01, l11);
    // ---- Q-function (iterative_ilqr.py:112-116), products written as nested fma chains
    const R G0 = fma(dt, V02, fma(b1, V01, b0 * V00)), G1 = fma(dt, V12, fma(b1, V11, b0 * V01));
    const R G2 = fma(dt, V22, fma(b1, V12, b0 * V02)), G3 = fma(dt, V23, fma(b1, V13, b0 * V03));
    const R H0 = dt * V03, H1 = dt * V13, H2 = dt * V23, H3 = dt * V33;
    const R X02 = fma(al, G0, fma(ga, G1, G2)), X03 = fma(be, G0, fma(de, G1, G3));  // Qux row 0 = [G0 G1 X02 X03]
    const R X12 = fma(al, H0, fma(ga, H1, H2)), X13 = fma(be, H0, fma(de, H1, H3));  // Qux row 1 = [H0 H1 X12 X13]
    const R W02 = fma(al, V00, fma(ga, V01, V02)), W12 = fma(al, V01, fma(ga, V11, V12));
    const R W22 = fma(al, V02, fma(ga, V12, V22)), W03 = fma(be, V00, fma(de, V01, V03));
    const R W13 = fma(be, V01, fma(de, V11, V13)), W23 = fma(be, V02, fma(de, V12, V23));
    const R W33 = fma(be, V03, fma(de, V13, V33));
    R Qx0, Qx1, Qx2, Qx3, Qu0, Qu1, Quu00, Quu01, Quu10, Quu11;
    R M00, M01, M02, M03, M11, M12, M13, M22, M23, M33;
    if (kQR) {
      const R d0 = xk[0] - p.xtarget[0], d1 = xk[1] - p.xtarget[1], d2 = xk[2] - p.xtarget[2],
                   d3 = xk[3] - p.xtarget[3];
      const R lx0 = (p.Q2[0] * d0 + p.Q2[1] * d1 + p.Q2[2] * d2 + p.Q2[3] * d3) + g0;
      const R lx1 = (p.Q2[1] * d0 + p.Q2[4] * d1 + p.Q2[5] * d2 + p.Q2[6] * d3) + g1;
      const R lx2 = p.Q2[2] * d0 + p.Q2[5] * d1 + p.Q2[7] * d2 + p.Q2[8] * d3;
      const R lx3 = p.Q2[3] * d0 + p.Q2[6] * d1 + p.Q2[8] * d2 + p.Q2[9] * d3;
      lu0 = (p.R2[0] * ua + p.R2[1] * ud) + lu0;
      lu1 = (p.R2[1] * ua + p.R2[2] * ud) + lu1;
      luu00 += p.R2[0];
      luu11 += p.R2[2];
      Qx0 = lx0 + Vx0;
      Qx1 = lx1 + Vx1;
      Qx2 = lx2 + fma(al, Vx0, fma(ga, Vx1, Vx2));
      Qx3 = lx3 + fma(be, Vx0, fma(de, Vx1, Vx3));
      Quu01 = fma(G3, dt, p.R2[1]);
      Quu10 = p.R2[1] + fma(H2, dt, fma(H1, b1, H0 * b0));
      M00 = (p.Q2[0] + l00) + V00;
      M01 = (p.Q2[1] + l01) + V01;
      M11 = (p.Q2[4] + l11) + V11;
      M02 = p.Q2[2] + W02;
      M03 = p.Q2[3] + W03;
      M12 = p.Q2[5] + W12;
      M13 = p.Q2[6] + W13;
      M22 = p.Q2[7] + fma(al, W02, fma(ga, W12, W22));
      M23 = p.Q2[8] + fma(al, W03, fma(ga, W13, W23));
      M33 = p.Q2[9] + fma(be, W03, fma(de, W13, W33));
    } else {  // matrix_Q = matrix_R = 0: l_x, l_xx live in the position block only, l_uu is diagonal
      Qx0 = g0 + Vx0;
      Qx1 = g1 + Vx1;
      Qx2 = fma(al, Vx0, fma(ga, Vx1, Vx2));
      Qx3 = fma(be, Vx0, fma(de, Vx1, Vx3));
      Quu01 = G3 * dt;
      Quu10 = fma(H2, dt, fma(H1, b1, H0 * b0));
      M00 = l00 + V00;
      M01 = l01 + V01;
      M11 = l11 + V11;
      M02 = W02;
      M03 = W03;
      M12 = W12;
      M13 = W13;
      M22 = fma(al, W02, fma(ga, W12, W22));
      M23 = fma(al, W03, fma(ga, W13, W23));
      M33 = fma(be, W03, fma(de, W13, W33));
    }
    Qu0 = fma(dt, Vx2, fma(b1, Vx1, fma(b0, Vx0, lu0)));
    Qu1 = fma(dt, Vx3, lu1);
    Quu00 = fma(G2, dt, fma(G1, b1, fma(G0, b0, luu00)));
    Quu11 = fma(H3, dt, luu11);
    // ---- regularised inverse and gains: iterative_ilqr.py:118-126
    R i00, i01, i11;
    quu_inverse(Quu00, (Quu01 + Quu10) / R(2.0), Quu11, c.lamb, i00, i01, i11);
    // The gains are kept with the opposite sign (k0 = -k[0], K00 = -K[0][0], ...): ten negation instructions per
    // stage less; the value-function update below is invariant under the flip and the forward pass subtracts,
    // so every result is bit-identical to the reference's sign convention.
    const R k0 = fma(i01, Qu1, i00 * Qu0), k1 = fma(i11, Qu1, i01 * Qu0);
    const R K00 = fma(i01, H0, i00 * G0), K01 = fma(i01, H1, i00 * G1);
    const R K02 = fma(i01, X12, i00 * X02), K03 = fma(i01, X13, i00 * X03);
    const R K10 = fma(i11, H0, i01 * G0), K11 = fma(i11, H1, i01 * G1);
    const R K12 = fma(i11, X12, i01 * X02), K13 = fma(i11, X13, i01 * X03);
    {
      const R gv[10] = {k0, k1, K00, K01, K02, K03, K10, K11, K12, K13};
      s.stg10(k, gv);
    }
    s.bwd_done(k);
    // ---- value function update with the unregularised Quu: iterative_ilqr.py:128-129
    const R P00 = fma(K10, Quu10, K00 * Quu00), P01 = fma(K10, Quu11, K00 * Quu01);
    const R P10 = fma(K11, Quu10, K01 * Quu00), P11 = fma(K11, Quu11, K01 * Quu01);
    const R P20 = fma(K12, Quu10, K02 * Quu00), P21 = fma(K12, Quu11, K02 * Quu01);
    const R P30 = fma(K13, Quu10, K03 * Quu00), P31 = fma(K13, Quu11, K03 * Quu01);
    Vx0 = fma(-P01, k1, fma(-P00, k0, Qx0));
    Vx1 = fma(-P11, k1, fma(-P10, k0, Qx1));
    Vx2 = fma(-P21, k1, fma(-P20, k0, Qx2));
    Vx3 = fma(-P31, k1, fma(-P30, k0, Qx3));
    V00 = fma(-P01, K10, fma(-P00, K00, M00));
    V01 = fma(-P01, K11, fma(-P00, K01, M01));
    V02 = fma(-P01, K12, fma(-P00, K02, M02));
    V03 = fma(-P01, K13, fma(-P00, K03, M03));
    V11 = fma(-P11, K11, fma(-P10, K01, M11));
    V12 = fma(-P11, K12, fma(-P10, K02, M12));
    V13 = fma(-P11, K13, fma(-P10, K03, M13));
    V22 = fma(-P21, K12, fma(-P20, K02, M22));
    V23 = fma(-P21, K13, fma(-P20, K03, M23));
    V33 = fma(-P31, K13, fma(-P30, K03, M33));
  }
}

// Full-step forward pass (iterative_ilqr.py:133-160).  Consumes the gains, leaves the trial trajectory in
// the new-slots, returns cost_new; Jtarget is the cost the reference would recompute for the trial
// trajectory at the top of its next iteration (stage cost about xtarget, iterative_ilqr.py:43-48).
template <bool kQR, int kN, class Scratch, typename R = typename Scratch::real>
__device__ __forceinline__ R forward_pass(const KParamsT<R>& p, const CandT<R>& c, const Scratch& s, R& Jtarget) {
  const int N = kN > 0 ? kN : p.N;  // kN > 0: horizon known at compile time, stage loops fully unrolled
  R x[4];
  s.load_x0(c, x);
  R cs, sn;
  s.load_cs0(c, cs, sn);
  R Jn = R(0.0), Jt = R(0.0);
  R g[10], gn[10];  // gains of the current stage; the next stage's are loaded one stage ahead
  s.fwd_begin(N);
  s.ldg10(0, g);
#pragma unroll
  for (int i = 0; i < 10; ++i) gn[i] = g[i];
#pragma unroll(kN > 0 ? kN : 2)  // two stages per trip: the one-stage-ahead operand registers need no rotation moves
  for (int k = 0; k < N; ++k) {
    s.fwd_wait(k);
    s.ldg10(k + 1 < N ? k + 1 : k, gn);  // clamped, not conditional: the prefetch must never leave the lane's strip
    R xo[4];
    if (k == 0) {
      s.load_x0(c, xo);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) xo[i] = s.ldcs(k - 1, 2 + i);
    }
    const R uo0 = s.ldcs(k, 0), uo1 = s.ldcs(k, 1);
    const R d0 = x[0] - xo[0], d1 = x[1] - xo[1], d2 = x[2] - xo[2], d3 = x[3] - xo[3];
    // u + k + K (x - xo) with the sign-flipped gains of backward_pass
    R ua = (uo0 - g[0]) - fma(g[5], d3, fma(g[4], d2, fma(g[3], d1, g[2] * d0)));
    R ud = (uo1 - g[1]) - fma(g[9], d3, fma(g[8], d2, fma(g[7], d1, g[6] * d0)));
    ua = clipd(ua, p.a_min, p.a_max);
    ud = clipd(ud, p.d_min, p.d_lim);
    if (kQR) {
      const R l_ctrl = quad2(p.Rm, ua, ud);
      Jn = Jn + quad4(p.Q, x[0] - c.xt[0], x[1] - c.xt[1], x[2] - c.xt[2], x[3] - c.xt[3]) + l_ctrl;
      Jt = Jt + quad4(p.Q, x[0] - p.xtarget[0], x[1] - p.xtarget[1], x[2] - p.xtarget[2], x[3] - p.xtarget[3]) + l_ctrl;
    }
    const R sd = x[2] * p.dt + ua * p.hdt2;
    x[0] = x[0] + cs * sd;
    x[1] = x[1] + sn * sd;
    x[2] = x[2] + ua * p.dt;
    x[3] = x[3] + ud * p.dt;
    sincos_fast(x[3], &sn, &cs);
    {
      const R tv[8] = {ua, ud, x[0], x[1], x[2], x[3], cs, sn};
      s.stg8(k, tv);
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) g[i] = gn[i];
    s.fwd_done(k);
  }
  R xt[4];
  s.load_xt(c, xt);
  const R T = terminal_cost_to(p, xt, x);
  Jtarget = Jt + T;
  return Jn + T;
}

// fabs((Jn - J) / J) < eps (iterative_ilqr.py:77), decided without the division whenever the division's rounding cannot
// change the answer.  With d = |Jn - J| and t = eps |J| (both rounded once): d < t (1 - 2^-50) implies the exact quotient
// lies below eps (1 - 2^-51), so its rounding is still below eps; d > t (1 + 2^-50) implies the exact quotient exceeds
// eps, and rounding is monotone.  In the band between (and for J = 0, infinities, NaN, where every comparison here is
// false) the reference's own expression decides.  Bit-identical decisions, one FP64 division less per iteration.
template <typename R>
__device__ __forceinline__ bool rel_change_below(R Jn, R J, R eps) {
  const R band = sizeof(R) == 8 ? R(8.881784197001252e-16) : R(4.76837158203125e-07);  // 2^-50, 2^-21
  const R dJ = Jn - J;
  const R d = fabs(dJ), t = eps * fabs(J);
  if (d < t * (R(1.0) - band)) return true;
  if (d > t * (R(1.0) + band)) return false;
  return fabs(dJ / J) < eps;
}

// One iLQR iteration with the accept / reject and lambda schedule of iterative_ilqr.py:74-84.
// Returns true when the candidate's loop ends.  Written so that every lane of a warp can execute it
// convergently (lanes without a live candidate pass live = false and only burn their own scratch): the
// tensor-memory scratch needs warp-aligned accesses.
template <bool kQR, int kN, class Scratch, typename R = typename Scratch::real>
__device__ __forceinline__ bool ilqr_iteration(const KParamsT<R>& p, CandT<R>& c, const Scratch& s, bool live = true) {
  backward_pass<kQR, kN>(p, c, s);
  s.fence_gains();
  // lambda of an accepted step, evaluated here so that the division's dependent chain drains under the forward pass
  const R lamb_accepted = c.lamb / p.lamb_factor;
  R Jt;
  const R Jn = forward_pass<kQR, kN>(p, c, s, Jt);
  s.fence_gains();
  const int N = kN > 0 ? kN : p.N;
  const bool accept = live && Jn < c.J;
  // (a streaming scratch moves whole warp-wide slots: the copy is skipped only when no lane of the warp accepted)
  if (Scratch::kStream ? __any_sync(0xffffffffu, accept) : (Scratch::kAligned || accept)) {
    // trial -> current; the trial of stage k+1 is loaded while stage k is stored
    R t[8], tn[8];
    s.cpy_begin(N);
    s.ldg8(0, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) tn[i] = t[i];
#pragma unroll 2
    for (int k = 0; k < N; ++k) {
      s.cpy_wait(k);
      s.ldg8(k + 1 < N ? k + 1 : k, tn);  // clamped prefetch (see forward_pass)
      if (accept) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s.stc(I2_CUR(k) + i, t[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = tn[i];
      s.cpy_done(k);
    }
  }
  if (!live) return false;
  c.iters++;
  bool done;
  if (accept) {
    c.lamb = lamb_accepted;
    done = rel_change_below(Jn, c.J, p.eps);
    c.J = kQR ? Jt : Jn;
  } else {
    c.lamb *= p.lamb_factor;
    done = c.lamb > p.max_lamb;
  }
  return done || c.iters >= p.max_iter;
}

// Relax-cost ladder (utils/base.py:427-437); `scale` = 10**it of the outer iteration.
template <typename R>
__device__ __forceinline__ double relax_cost(const KParamsT<R>& p, const R* xN, const R* xt, double qfun, double scale) {
  const double e0 = (double)xN[0] - (double)xt[0], e1 = (double)xN[1] - (double)xt[1], e2 = (double)xN[2] - (double)xt[2],
               e3 = (double)xN[3] - (double)xt[3];
  const double r = sqrt(e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3);
  for (int i = 1; i <= p.max_relax; ++i) {
    if (r <= 80.0 * i / scale) return qfun + p.N + 100 * i;
    if (r > 80.0 * p.max_relax / scale) return INFINITY;
  }
  return NAN;
}

// Result in the reference's array layout: xvar (4, N+1), uvar (2, N), time on the last axis.
template <int kN, class Scratch, typename R = typename Scratch::real>
__device__ __forceinline__ void store_candidate(const KParamsT<R>& p, const CandT<R>& c, const Scratch& s, double* xvar,
                                                double* uvar) {
  const int N = kN > 0 ? kN : p.N;  // kN > 0: horizon known at compile time, stage loops fully unrolled
  if (xvar) {
    R x0[4];
    s.load_x0(c, x0);
#pragma unroll
    for (int i = 0; i < 4; ++i) xvar[i * (N + 1)] = x0[i];
  #pragma unroll(kN > 0 ? kN : 1)
  for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int i = 0; i < 4; ++i) xvar[i * (N + 1) + k + 1] = s.ldc(I2_CUR(k) + 2 + i);
    }
  }
  if (uvar) {
  #pragma unroll(kN > 0 ? kN : 1)
  for (int k = 0; k < N; ++k) {
      uvar[k] = s.ldc(I2_CUR(k) + 0);
      uvar[N + k] = s.ldc(I2_CUR(k) + 1);
    }
  }
}

template <typename R>
__device__ __forceinline__ void load_obstacle(const KParamsT<R>& p, CandT<R>& c, const DevObstacle* o) {
  if (o != nullptr && o->present) {
    c.ox = R(o->x);
    c.oy = R(o->y);
    c.oP0 = R(1.0 / (o->width * o->width));
    c.oP1 = R(1.0 / (o->height * o->height));
    c.osx = R(o->moving_option == 2 ? o->spd : 0.0);
    c.osy = R(o->moving_option == 1 ? o->spd : 0.0);
    c.og = p.og;
    c.oh = p.oh;
  } else {
    c.ox = c.oy = c.oP0 = c.oP1 = c.osx = c.osy = c.og = c.oh = R(0.0);
  }
  c.ex = nullptr;
  c.n_ex = 0;
}

}  // namespace i2lqr
