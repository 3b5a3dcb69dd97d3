// kernels.cuh -- the sm_100a kernels of the i2LQR control solve.
//
//   K1  knn_kernel      safe-set K-nearest selection   (iLqr.select_close_ss, utils/base.py:332-341)
//   K2  solve_kernel    persistent batched iLQR, one lane per candidate / per lambda chain
//                       (control/iterative_ilqr.py:7-160 + the candidate loop of utils/base.py:403-455)
//   K3  winner_kernel   lexicographic winner + result gather (utils/base.py:462-478)
//   K4  plant_kernel    batched closed-loop episode step (utils/base.py:157-213, 708-726)
#pragma once
#include <stdint.h>

#include <type_traits>

#include "chain_core.cuh"
#include "ilqr_core.cuh"

namespace i2lqr {

constexpr int kOuter = 3;  // calc_input's outer iterations (utils/base.py:384, break at :472-478)

// Device-resident safe sets: per episode a ring of L lap slots, each SoA [4][Tcap].
struct SafeSet {
  double* ss;      // [E][L][4][Tcap]
  int* lap_len;    // [E][L]
  int* lap_count;  // [E]  laps ever added; absolute lap a lives in slot a % L
  int L, Tcap;
  __device__ __forceinline__ const double* lap(int e, int slot) const {
    return ss + ((size_t)e * L + slot) * 4 * (size_t)Tcap;
  }
};

struct Work {
  int mode;                    // 0 = flat candidates, 1 = lambda chains of a controller step
  long long n_items;           // candidates (mode 0) or chains = E*L (mode 1)
  unsigned long long* counter; // work-fetch counter (zeroed before launch)
  unsigned long long* stats;   // [0] += executed iLQR iterations, [1] += finished candidate solves
  double* gscratch;            // global scratch slab when the horizon does not fit shared memory
  long long gstride;
  // ---- mode 0
  const double* x0;            // (B,4)
  const double* x_term;        // (B,4)
  const double* lamb_in;       // (B) or null (= param.lamb)
  const DevObstacle* obs;      // n_obs entries
  long long n_obs;             // 0, 1 or B
  const double* u_init;        // (B,2,N) warm start or null
  double* uvar;                // (B,2,N)
  double* xvar;                // (B,4,N+1)
  double* lamb_out;            // (B)
  int* iters;                  // (B)
  // ---- mode 0, iteration-sliced passes (see solve_kernel)
  int pass;                    // 0 = fresh candidates, > 0 = resume parked ones
  int iter_cap;                // park a candidate that reaches this many iterations unfinished (0 = never)
  long long park_capacity;     // entries the park buffers hold (a lane that finds them full just keeps iterating)
  const unsigned long long* resume_count;  // device: entries the previous pass tried to park
  const int* resume_list;      // candidate ids parked by the previous pass
  const double* resume_state;  // (capacity, 8N+4): current trajectory slots, lambda, J, iterations
  unsigned long long* park_count;          // device: entries this pass parks
  int* park_list;
  double* park_state;
  // ---- mode 1
  int it;                      // outer iteration 0..2
  double scale;                // 10**it
  SafeSet set;
  const double* x;             // (E,4) current states
  const DevObstacle* ep_obs;   // (E)
  const DevObstacle* ep_obs_extra;  // (E, kMaxExtraObstacles) further obstacles per episode, or null
  const int* ep_n_extra;       // (E) how many of them are in use
  const int* active;           // (E) or null: episodes that solve this step
  const int* ep_list;          // compacted list of the episodes that solve this step, or null (= all n_items / L episodes)
  const int* n_ep;             // device: entries of ep_list
  const long long* sel_idx;    // (E,3,L,K)
  const int* sel_cnt;          // (E,L)
  double* cand_x;              // (E,L,K,4,N+1)
  double* cand_u;              // (E,L,K,2,N)
  double* cand_cost;           // (E,3,L,K)
  double* cand_lamb;           // (E,3,L,K)
  int* cand_iters;             // (E,3,L,K)
};

// ------------------------------------------------------------------------------------------------ K2
// Persistent kernel, one lane per candidate.  Every lane runs a small state machine -- take a candidate,
// INIT it, RUN one iLQR iteration per trip of the warp loop, finalise -- so a warp always executes the same
// code (one iteration) on up to 32 different candidates regardless of how many iterations each needs
// (1..150).  kMode 0: flat batch; the NEXT candidate of every lane (index from a warp-aggregated atomic and
// its inputs) is prefetched into registers while the current one iterates, so the global-load latency of
// the few lanes that switch candidates in a trip never stalls the warp.  kMode 1: lambda chains of a
// controller step; a lane walks the K candidates of its (episode, lap) in rank order carrying lambda
// (utils/base.py:393,414).
__device__ __forceinline__ void prefetch_l2(const void* q) { asm volatile("prefetch.global.L2 [%0];" ::"l"(q)); }
enum : int { ST_NEED = 0, ST_INIT = 1, ST_RUN = 2, ST_DONE = 3, ST_FIN = 4, ST_PARK = 5 };
#ifndef I2LQR_SERVICE_TRIPS
#define I2LQR_SERVICE_TRIPS 2  // flat batches: candidate switches every this many trips of the warp loop (see solve_kernel)
#endif
constexpr int kSolveThreads = 32;    // block size of the memory-backed variants (one warp: its SM slot frees as soon as it is done)
constexpr int kTmemThreads = 128;    // block size of the tensor-memory variant (a CTA's TMEM allocation spans four lane quarters)
constexpr int kTmemMaxHorizon = 25;  // 20 32-bit TMEM columns per stage: 128 columns up to N = 6, 256 up to 12, 512 up to 25
__host__ __device__ __forceinline__ int tmem_columns(int N) { return 20 * N <= 128 ? 128 : (20 * N <= 256 ? 256 : 512); }

// kScratch: 0 = shared memory, 1 = global slab, 2 = hybrid (trajectory in shared memory, gains in the slab),
//           3 = trajectory in shared memory, gains in tensor memory (12 resident warps per SM instead of 8),
//           4 = global memory streamed through a shared-memory ring by bulk asynchronous copies (long horizons)
constexpr int kStreamRing = 8;  // ring slots per warp (4 KB each): the loads of the next 6-7 stages are in flight
template <typename R, int kScratch, int kThreads>
struct ScratchSelect {
  using type = typename std::conditional<kScratch == 0, SmemScratch<R, kThreads>,
               typename std::conditional<kScratch == 1, GlobalScratch<R>,
               typename std::conditional<kScratch == 2, HybridScratch<R, kThreads>,
               typename std::conditional<kScratch == 3, TmemScratch<kThreads>,
               typename std::conditional<kScratch == 4, StreamScratch<kStreamRing>,
               typename std::conditional<kScratch == 6, TmemLeanScratch<kThreads>, StreamScratch<4>>::type>::type>::type>::type>::type>::type;
};  // (kScratch 5 = the streaming placement with a 4-slot ring: half the shared memory per warp, twice the warps per SM;
    //  kScratch 6 = the tensor-memory placement with the candidate's cold state parked too: 128 registers, four CTAs per SM)
constexpr int kLeanMaxHorizon = 6;  // 20 N + 8 tensor-memory columns within a CTA's 128
__host__ __device__ __forceinline__ size_t stream_smem_bytes(int ring = kStreamRing) { return (size_t)ring * 16 * 32 * sizeof(double) + ring * 8; }
__host__ __device__ __forceinline__ size_t stream_warp_doubles(int N) { return (size_t)N * 16 * 32; }

// R = double is the product path (results within 1e-9 of the reference); R = float is the optional FP32
// build of BASELINE config 4 (inputs / outputs stay double, everything in between is single precision).
template <typename R, bool kQR, int kScratch, int kMode, int kThreads>
__global__ void __launch_bounds__(kThreads, kScratch == 3 ? 3 : (kScratch == 6 ? 4 : (kScratch == 5 ? 12 : 1)))
solve_kernel(const __grid_constant__ KParamsT<R> p, const __grid_constant__ Work w) {
  extern __shared__ double smem_raw[];
  R* smem = reinterpret_cast<R*>(smem_raw);
  using ScratchT = typename ScratchSelect<R, kScratch, kThreads>::type;
  ScratchT s;
  __shared__ uint32_t tmem_base;
  constexpr bool kTmem = kScratch == 3 || kScratch == 6;
  constexpr bool kLean = kScratch == 6;
  if constexpr (kTmem) {
    if (threadIdx.x < 32) {  // one warp allocates the CTA's tensor-memory columns
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(&tmem_base)),
                   "r"(tmem_columns(p.N))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (kLean) {
      // (the thread index goes through an opaque move: otherwise the lane's shared-memory base is re-derived from
      //  SR_TID.X -- six instructions -- in every stage of every loop rather than kept in its register)
      unsigned tid = threadIdx.x;
      asm volatile("mov.u32 %0, %0;" : "+r"(tid));
      s.bind_lean(smem + tid, tmem_base + (((tid >> 5) * 32u) << 16), p.N);
      CandT<R> z = {};
      s.park_start(z);
      s.park_target(z, true);
    } else {
      s.bind_tmem(smem + threadIdx.x, tmem_base + (((threadIdx.x >> 5) * 32u) << 16));
    }
    for (int e = 0; e < 8 * p.N; ++e) s.stc(e, R(0.0));  // lanes without a live candidate iterate on zeros
  } else if constexpr (kScratch == 4 || kScratch == 5) {
    s.bind_stream(smem_raw, w.gscratch + (size_t)blockIdx.x * stream_warp_doubles(p.N), threadIdx.x & 31, p.N);
    for (int e = 0; e < 16 * p.N; ++e) s.gcur[(size_t)e * 32 + (threadIdx.x & 31)] = 0.0;  // lanes without a live candidate iterate on zeros
  } else {
    s.bind(smem + threadIdx.x, reinterpret_cast<R*>(w.gscratch) + (size_t)blockIdx.x * blockDim.x + threadIdx.x, w.gstride, p.N);
  }
  constexpr int kN = 0;  // stage loops stay rolled: the fully unrolled N=6 body measured 1.4x slower (I-cache, spills)
  const unsigned lane = threadIdx.x & 31u;
  const int N = p.N;
  CandT<R> c = {};
  int state = ST_NEED;
  // (the lean placement counts in 32 bits: a launch serves fewer than 2^31 candidates and a lane fewer than 2^32 iterations)
  using index_t = typename std::conditional<kScratch == 6, int, long long>::type;
  using count_t = typename std::conditional<kScratch == 6, unsigned, unsigned long long>::type;
  index_t item = 0;
  count_t my_iters = 0, my_solves = 0;

  if constexpr (kMode == 0) {
    // ---- prefetched next candidate
    bool have_next = false, exhausted = false;
    index_t n_item = 0;
    long long n_total = w.n_items;
    if (w.pass > 0) {
      n_total = (long long)*w.resume_count;
      if (n_total > w.park_capacity) n_total = w.park_capacity;
    }
    const int park_words = 8 * N + 4;
    index_t n_idx = 0, idx = 0;  // position in the pass's work list (== candidate id in pass 0)
    // (the lean placement keeps no prefetched inputs in registers: it asks for the lines in L2 and loads them when the
    //  lane takes the candidate, a service trip or more later)
    double n_x0[4], n_xt[4], n_lamb = (double)p.lamb0;
    DevObstacle n_obs;
    n_obs.present = 0;
    // Candidate switches are batched: a lane that finishes (or reaches the pass's iteration cap) only flags itself and
    // idles; every kServiceTrips-th trip of the warp loop is a "service trip" in which all flagged lanes store their
    // results / park, and lanes without a candidate take the next one.  The switch code runs with only the lanes that
    // switch active -- 2.7 of 32 when it ran every trip, a quarter of the kernel's time at 10 % lane efficiency
    // (profiles/r01_solve_kernel_v6_summary.txt) -- so running it every other trip with twice the lanes halves that cost;
    // the price is a finished lane idling half a trip on average (4 % of an 11.7-iteration candidate).  Scheduling
    // only: every candidate's arithmetic is unchanged.
    constexpr int kServiceTrips = I2LQR_SERVICE_TRIPS;
    int trip = 0;
    bool no_park = false;
    for (;;) {
      const bool service = kServiceTrips == 1 || trip == 0;  // warp-uniform
      trip = trip + 1 == kServiceTrips ? 0 : trip + 1;
      if (service) {
        // (z) lanes flagged in the previous trips: results out, or park
        if constexpr (ScratchT::kStream) {  // long horizons: the warp writes each finished lane's trajectory together
          __syncwarp();
          unsigned fm = __ballot_sync(0xffffffffu, state == ST_FIN);
          while (fm) {
            const int j = __ffs(fm) - 1;
            fm &= fm - 1;
            const long long it_j = __shfl_sync(0xffffffffu, item, j);
            double x0j[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) x0j[i] = __shfl_sync(0xffffffffu, (double)c.x0[i], j);
            s.store_lane(j, x0j, w.xvar ? w.xvar + it_j * 4 * (N + 1) : nullptr, w.uvar ? w.uvar + it_j * 2 * N : nullptr);
          }
          if (state == ST_FIN) {
            if (w.lamb_out) w.lamb_out[item] = (double)c.lamb;
            if (w.iters) w.iters[item] = c.iters;
            state = ST_NEED;
          }
        } else if (state == ST_FIN) {
          store_candidate<kN>(p, c, s, w.xvar ? w.xvar + (long long)item * 4 * (N + 1) : nullptr, w.uvar ? w.uvar + (long long)item * 2 * N : nullptr);
          if (w.lamb_out) w.lamb_out[item] = (double)c.lamb;
          if (w.iters) w.iters[item] = c.iters;
          state = ST_NEED;
        }
        // Iteration slicing: a candidate that is still unfinished at the pass's cap is parked (trajectory,
        // lambda, J, count) and its lane takes the next one.  The following pass resumes the parked ones packed
        // into full warps, so the few-percent long tail (up to max_ilqr_iter = 150 iterations against a mean
        // of ~12) no longer runs as hundreds of trips of nearly empty warps.
        const bool park = state == ST_PARK;
        const unsigned parkers = __ballot_sync(0xffffffffu, park);
        if (parkers) {
          unsigned long long pbase = 0;
          const int leader = __ffs(parkers) - 1;
          if ((int)lane == leader) pbase = atomicAdd(w.park_count, (unsigned long long)__popc(parkers));
          pbase = __shfl_sync(0xffffffffu, pbase, leader);
          const long long pslot = (long long)pbase + __popc(parkers & ((1u << lane) - 1u));
          if (park) {
            if (pslot < w.park_capacity) {
              double* ps = w.park_state + (size_t)pslot * park_words;
              for (int e = 0; e < 8 * N; ++e) ps[e] = (double)s.ldc(e);
              ps[8 * N] = (double)c.lamb;
              ps[8 * N + 1] = (double)c.J;
              ps[8 * N + 2] = (double)c.iters;
              w.park_list[pslot] = (int)item;
              state = ST_NEED;
            } else {
              state = ST_RUN;  // park buffers full: this candidate runs to its end in this pass
              no_park = true;
            }
          }
        }
        // (a) lanes without a pending candidate fetch an index and issue its loads (consumed trips later)
        const unsigned want = __ballot_sync(0xffffffffu, !have_next && !exhausted);
        if (want) {
          unsigned long long base = 0;
          const int leader = __ffs(want) - 1;
          if ((int)lane == leader) base = atomicAdd(w.counter, (unsigned long long)__popc(want));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (!have_next && !exhausted) {
            const long long fetched = (long long)(base + __popc(want & ((1u << lane) - 1u)));
            n_idx = (index_t)fetched;
            if (fetched >= n_total) {
              exhausted = true;
            } else {
              n_item = w.pass > 0 ? (index_t)w.resume_list[n_idx] : n_idx;
              have_next = true;
              if constexpr (kLean) {
                prefetch_l2(w.x0 + (long long)n_item * 4);
                prefetch_l2(w.x_term + (long long)n_item * 4);
                if (w.lamb_in) prefetch_l2(w.lamb_in + n_item);
                if (w.n_obs > 1) prefetch_l2(w.obs + n_item);
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  n_x0[i] = w.x0[n_item * 4 + i];
                  n_xt[i] = w.x_term[n_item * 4 + i];
                }
                if (w.lamb_in) n_lamb = w.lamb_in[n_item];
                if (w.n_obs != 0) n_obs = w.obs[w.n_obs == 1 ? 0 : n_item];
              }
            }
          }
        }
        // (b) lanes that need a candidate take the prefetched one
        bool took = false;
        if (state == ST_NEED) {
          if (have_next) {
            item = n_item;
            idx = n_idx;
            if constexpr (kLean) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                n_x0[i] = w.x0[(long long)item * 4 + i];
                n_xt[i] = w.x_term[(long long)item * 4 + i];
              }
              if (w.lamb_in) n_lamb = w.lamb_in[item];
              if (w.n_obs != 0) n_obs = w.obs[w.n_obs == 1 ? 0 : item];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              c.x0[i] = R(n_x0[i]);
              c.xt[i] = R(n_xt[i]);
            }
            c.lamb = R(n_lamb);
            load_obstacle(p, c, w.n_obs != 0 ? &n_obs : nullptr);
            sincos_fast(c.x0[3], &c.s0, &c.c0);
            took = true;
            if (w.pass == 0) {
              init_candidate<kQR, kN>(p, c, s, w.u_init != nullptr ? w.u_init + (long long)item * 2 * N : nullptr);
            } else {  // resume a candidate parked by the previous pass
              const double* ps = w.resume_state + (size_t)idx * park_words;
              for (int e = 0; e < 8 * N; ++e) s.stc(e, R(ps[e]));
              c.lamb = R(ps[8 * N]);
              c.J = R(ps[8 * N + 1]);
              c.iters = (int)ps[8 * N + 2];
            }
            have_next = false;
            no_park = false;
            state = ST_RUN;
            if constexpr (kLean) s.park_start(c);
          } else if (exhausted) {
            state = ST_DONE;
          }
        }
        if constexpr (kLean) {
          __syncwarp();
          s.park_target(c, took);  // warp-collective tensor-memory store
        }
      }
      if (__all_sync(0xffffffffu, state == ST_DONE)) break;
      // Every lane of the warp is still in this loop (finished lanes idle until the collective break), so the
      // full-mask vote above also reconverges the warp before the iteration; without that the freshly
      // initialised lanes run the whole iteration as a separate divergent pass (measured: 15.5 of 32
      // lanes active per issued instruction).
      const bool live = state == ST_RUN;
      bool finished;
      if constexpr (ScratchT::kAligned) {
        finished = ilqr_iteration<kQR, kN>(p, c, s, live);  // warp-collective scratch: all 32 lanes go through
      } else {
        finished = live && ilqr_iteration<kQR, kN>(p, c, s);
      }
      if (live) {
        if (finished) {
          my_solves += 1;
          state = ST_FIN;
        } else if (w.iter_cap > 0 && c.iters >= w.iter_cap && !no_park) {
          state = ST_PARK;
        }
        my_iters += 1;
      }
    }
  } else {
    int rank = 0, cnt = 0;  // position in the chain
    int ep = 0, lrel = 0, slot = 0, T = 0;
    const double* lap = nullptr;
    long long jsel = 0;
    for (;;) {
      const unsigned want = __ballot_sync(0xffffffffu, state == ST_NEED);
      if (want) {
        unsigned long long base = 0;
        const int leader = __ffs(want) - 1;
        if ((int)lane == leader) base = atomicAdd(w.counter, (unsigned long long)__popc(want));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (state == ST_NEED) {
          item = (long long)(base + __popc(want & ((1u << lane) - 1u)));
          if (item >= w.n_items) {
            state = ST_DONE;
          } else {
            ep = (int)(item / w.set.L);
            lrel = (int)(item % w.set.L);
            cnt = (w.active == nullptr || w.active[ep]) ? w.sel_cnt[(size_t)ep * w.set.L + lrel] : 0;
            if (cnt > 0) {  // else: unused lap slot or inactive episode, take another item next trip
              const int count = w.set.lap_count[ep];
              const int min_iter = count - w.set.L > 0 ? count - w.set.L : 0;
              slot = (min_iter + lrel) % w.set.L;
              T = w.set.lap_len[(size_t)ep * w.set.L + slot];
              lap = w.set.lap(ep, slot);
              rank = 0;
              c.lamb = p.lamb0;  // utils/base.py:393
              load_obstacle(p, c, w.ep_obs + ep);
#pragma unroll
              for (int i = 0; i < 4; ++i) c.x0[i] = R(w.x[(size_t)ep * 4 + i]);
              sincos_fast(c.x0[3], &c.s0, &c.c0);
              state = ST_INIT;
            }
          }
        }
      }
      if (__all_sync(0xffffffffu, state == ST_DONE)) break;
      if (state == ST_INIT) {
        jsel = w.sel_idx[(((size_t)ep * kOuter + w.it) * w.set.L + lrel) * p.K + rank];
#pragma unroll
        for (int i = 0; i < 4; ++i) c.xt[i] = R(lap[(size_t)i * w.set.Tcap + jsel]);
        init_candidate<kQR, kN>(p, c, s, nullptr);
        state = ST_RUN;
      }
      __syncwarp();  // reconverge before the iteration (see kMode 0)
      if (state == ST_RUN) {
        if (ilqr_iteration<kQR, kN>(p, c, s)) {
          my_iters += (unsigned long long)c.iters;
          my_solves += 1;
          const size_t cslot = ((size_t)ep * w.set.L + lrel) * p.K + rank;
          store_candidate<kN>(p, c, s, w.cand_x + cslot * 4 * (N + 1), w.cand_u + cslot * 2 * N);
          R xN[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xN[i] = s.ldc(I2_CUR(N - 1) + 2 + i);
          const size_t tslot = (((size_t)ep * kOuter + w.it) * w.set.L + lrel) * p.K + rank;
          w.cand_cost[tslot] = relax_cost(p, xN, c.xt, (double)(T - 1 - jsel), w.scale);
          w.cand_lamb[tslot] = (double)c.lamb;
          w.cand_iters[tslot] = c.iters;
          ++rank;
          state = rank < cnt ? ST_INIT : ST_NEED;  // lambda is carried to the lap's next candidate
        }
      }
    }
  }
  // ---- statistics: one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    my_iters += __shfl_xor_sync(0xffffffffu, my_iters, o);
    my_solves += __shfl_xor_sync(0xffffffffu, my_solves, o);
  }
  if (lane == 0 && w.stats != nullptr && (my_solves | my_iters)) {
    atomicAdd(w.stats + 0, my_iters);
    atomicAdd(w.stats + 1, my_solves);
  }
  if constexpr (kTmem) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_columns(p.N)) : "memory");
  }
}

// ------------------------------------------------------------------------------- K2, controller-step form
// Lambda chains with speculative Levenberg-Marquardt rungs.  The K candidates of one (episode, lap) are a
// sequential chain (lambda is carried, utils/base.py:393,414-426), and inside a candidate every rejected
// iteration repeats the backward and forward pass from the SAME trajectory with lambda * 10
// (iterative_ilqr.py:74-84; no line search) -- half of all iterations of a closed-loop run are
// rejects.  kRungs adjacent lanes therefore own one chain: all hold the same current trajectory, lane r
// evaluates the iteration with lambda * 10^r (the exact value the reference would reach after r rejects), and
// the group then accepts in the reference's order -- the first rung whose trial cost is below the current cost
// wins, the rungs before it count as the rejected iterations they are, the rungs after it are discarded.  One
// trip of the warp loop thus consumes 1..kRungs reference iterations; results are bit-identical to the
// one-lane chain (same code, same operands).  A chain is a static work item: lane group g of the grid owns
// chain g.
template <typename R, bool kQR, int kScratch, int kRungs>
__global__ void __launch_bounds__(kSolveThreads, 1)
chain_kernel(const __grid_constant__ KParamsT<R> p, const __grid_constant__ Work w) {
  extern __shared__ double smem_raw[];
  R* smem = reinterpret_cast<R*>(smem_raw);
  using ScratchT = typename ScratchSelect<R, kScratch, kSolveThreads>::type;
  static_assert(kScratch != 3, "the controller-step kernel keeps its scratch in memory");
  ScratchT s;
  s.bind(smem + threadIdx.x, reinterpret_cast<R*>(w.gscratch) + (size_t)blockIdx.x * blockDim.x + threadIdx.x, w.gstride, p.N);
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int rung = lane % kRungs;
  const int gbase_lane = lane - rung;  // first lane of this chain's group
  constexpr int kN = 0;
  const int N = p.N;
  const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kRungs;
  CandT<R> c = {};
  int state = ST_DONE;
  int rank = 0, cnt = 0, ep = 0, lrel = 0, T = 0;
  const double* lap = nullptr;
  long long jsel = 0;
  unsigned long long my_iters = 0, my_solves = 0;
  if (item < (w.n_ep != nullptr ? (long long)*w.n_ep * w.set.L : w.n_items)) {
    ep = (int)(item / w.set.L);
    if (w.ep_list != nullptr) ep = w.ep_list[ep];  // batched episodes: only the ones that solve this step
    lrel = (int)(item % w.set.L);
    cnt = (w.active == nullptr || w.active[ep]) ? w.sel_cnt[(size_t)ep * w.set.L + lrel] : 0;
    if (cnt > 0) {
      const int count = w.set.lap_count[ep];
      const int min_iter = count - w.set.L > 0 ? count - w.set.L : 0;
      const int slot = (min_iter + lrel) % w.set.L;
      T = w.set.lap_len[(size_t)ep * w.set.L + slot];
      lap = w.set.lap(ep, slot);
      c.lamb = p.lamb0;  // utils/base.py:393
      load_obstacle(p, c, w.ep_obs + ep);
#pragma unroll
      for (int i = 0; i < 4; ++i) c.x0[i] = R(w.x[(size_t)ep * 4 + i]);
      sincos_fast(c.x0[3], &c.s0, &c.c0);
      state = ST_INIT;
    }
  }
  for (;;) {
    if (__all_sync(kFull, state == ST_DONE)) break;
    if (state == ST_INIT) {
      jsel = w.sel_idx[(((size_t)ep * kOuter + w.it) * w.set.L + lrel) * p.K + rank];
#pragma unroll
      for (int i = 0; i < 4; ++i) c.xt[i] = R(lap[(size_t)i * w.set.Tcap + jsel]);
      init_candidate<kQR, kN>(p, c, s, nullptr);  // every rung keeps its own copy of the current trajectory
      state = ST_RUN;
    }
    __syncwarp();
    const bool run = state == ST_RUN;
    R Jn = R(0.0), Jt = R(0.0);
    if (run) {
      const R lamb_chain = c.lamb;
      for (int r = 0; r < rung; ++r) c.lamb *= p.lamb_factor;  // the value r rejects lead to, same operations
      backward_pass<kQR, kN>(p, c, s);
      Jn = forward_pass<kQR, kN>(p, c, s, Jt);
      c.lamb = lamb_chain;
    }
    __syncwarp();  // trial trajectories are read across lanes below
    const unsigned acc_all = __ballot_sync(kFull, run && Jn < c.J);
    const unsigned acc = (acc_all >> gbase_lane) & ((1u << kRungs) - 1u);
    // ---- accept in the reference's order (iterative_ilqr.py:74-84); every lane of the group computes the same
    int winner = -1;
    bool fin = false;
    int consumed = 0;
    R lamb = c.lamb;
    if (run) {
#pragma unroll
      for (int r = 0; r < kRungs; ++r) {
        if (winner >= 0 || fin) break;
        if (c.iters + consumed >= p.max_iter) {
          fin = true;
          break;
        }
        ++consumed;
        if ((acc >> r) & 1u) {
          winner = r;
          lamb /= p.lamb_factor;
        } else {
          lamb *= p.lamb_factor;
          if (lamb > p.max_lamb) fin = true;
        }
      }
    }
    const int src = winner >= 0 ? gbase_lane + winner : lane;
    const R Jn_w = __shfl_sync(kFull, Jn, src), Jt_w = __shfl_sync(kFull, Jt, src);
    if (run) {
      if (winner >= 0) {
        const int delta = winner - rung;
        for (int k = 0; k < N; ++k) {
#pragma unroll
          for (int i = 0; i < 8; ++i) s.stc(I2_CUR(k) + i, s.ldg_lane(delta, k * 10 + i));
        }
        if (rel_change_below(Jn_w, c.J, p.eps)) fin = true;
        c.J = kQR ? Jt_w : Jn_w;
      }
      c.lamb = lamb;
      c.iters += consumed;
      if (c.iters >= p.max_iter) fin = true;
    }
    __syncwarp();  // the winner's trial slots are overwritten by the next trip only after everyone copied them
    if (run && fin) {
      const size_t cslot = ((size_t)ep * w.set.L + lrel) * p.K + rank;
      const size_t tslot = (((size_t)ep * kOuter + w.it) * w.set.L + lrel) * p.K + rank;
      if (rung == 0) {
        my_iters += (unsigned long long)c.iters;
        my_solves += 1;
        store_candidate<kN>(p, c, s, w.cand_x + cslot * 4 * (N + 1), nullptr);
        R xN[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xN[i] = s.ldc(I2_CUR(N - 1) + 2 + i);
        w.cand_cost[tslot] = relax_cost(p, xN, c.xt, (double)(T - 1 - jsel), w.scale);
        w.cand_lamb[tslot] = (double)c.lamb;
        w.cand_iters[tslot] = c.iters;
      }
      if (rung == kRungs - 1) store_candidate<kN>(p, c, s, nullptr, w.cand_u + cslot * 2 * N);
      ++rank;
      state = rank < cnt ? ST_INIT : ST_DONE;  // lambda is carried to the lap's next candidate
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_iters += __shfl_xor_sync(kFull, my_iters, o);
    my_solves += __shfl_xor_sync(kFull, my_solves, o);
  }
  if (lane == 0 && w.stats != nullptr && (my_solves | my_iters)) {
    atomicAdd(w.stats + 0, my_iters);
    atomicAdd(w.stats + 1, my_solves);
  }
}

// Latency-oriented form of chain_kernel for horizons whose working set fits one warp's shared memory
// (chain_core.cuh): the group of kRungs lanes shares one current trajectory (pointer switch instead of a copy on
// accept) and the lambda-independent part of the backward pass (one task per stage, dealt to the group's lanes),
// every rung then runs the Riccati recursion and the skewed forward pass with its own lambda.  Same arithmetic
// and same accept order as chain_kernel.  A chain owns kGroup >= kRungs adjacent lanes: the first kRungs are the
// rungs, all kGroup share the stage tasks.  With few chains (one controller, config 2) kGroup = 32 gives every
// chain a warp, hence an SM sub-partition, of its own and one task per lane; with many chains (batched
// episodes) kGroup shrinks so that the chains still fit one warp per sub-partition.
template <typename R, bool kQR, int kRungs, int kGroup>
__global__ void __launch_bounds__(kSolveThreads, 1)
chain_lat_kernel(const __grid_constant__ KParamsT<R> p, const __grid_constant__ Work w) {
  extern __shared__ double smem_raw[];
  R* smem = reinterpret_cast<R*>(smem_raw);
  constexpr int S = kSolveThreads;
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int N = p.N;
  GroupSolver<R, kQR, kRungs, kGroup, S> gs;
  gs.bind(smem, lane, N);
  const int rung = gs.rung;
  CandT<R>& c = gs.c;
  // A chain is a work item.  Group g of the grid starts with chain g; when the launch has more chains than lane groups
  // (batched episodes: the grid is capped at what is resident at once) a group that finishes its chain takes the next
  // unassigned one from `w.counter`, so the chains' very different lengths average out instead of adding up per wave.
  const long long n_groups = (long long)gridDim.x * (blockDim.x / kGroup);
  const long long n_chains = w.n_ep != nullptr ? (long long)*w.n_ep * w.set.L : w.n_items;
  long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kGroup;
  int state = ST_DONE;
  int rank = 0, cnt = 0, ep = 0, lrel = 0, T = 0;
  const double* lap = nullptr;
  long long jsel = 0;
  unsigned long long my_iters = 0, my_solves = 0;
  auto take = [&]() {  // binds the group to chain `item` (all lanes of the group hold the same item)
    state = ST_DONE;
    if (item >= n_chains) return false;
    ep = (int)(item / w.set.L);
    if (w.ep_list != nullptr) ep = w.ep_list[ep];  // batched episodes: only the ones that solve this step
    lrel = (int)(item % w.set.L);
    cnt = (w.active == nullptr || w.active[ep]) ? w.sel_cnt[(size_t)ep * w.set.L + lrel] : 0;
    rank = 0;
    if (cnt > 0) {
      const int count = w.set.lap_count[ep];
      const int min_iter = count - w.set.L > 0 ? count - w.set.L : 0;
      const int slot = (min_iter + lrel) % w.set.L;
      T = w.set.lap_len[(size_t)ep * w.set.L + slot];
      lap = w.set.lap(ep, slot);
      c.lamb = p.lamb0;  // utils/base.py:393
      load_obstacle(p, c, w.ep_obs + ep);
      if (w.ep_n_extra != nullptr) {
        c.ex = w.ep_obs_extra + (size_t)ep * kMaxExtraObstacles;
        c.n_ex = w.ep_n_extra[ep];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) c.x0[i] = R(w.x[(size_t)ep * 4 + i]);
      sincos_fast(c.x0[3], &c.s0, &c.c0);
      state = ST_INIT;
    }
    return true;
  };
  // next chain of this group: the static one first, then from the shared counter (which starts at n_groups);
  // empty chains (unused lap slots) are skipped on the spot
  auto next_chain = [&](bool first) {
    for (;;) {
      if (!first) {
        if (w.counter == nullptr || n_chains <= n_groups) {
          state = ST_DONE;
          return;
        }
        unsigned long long nx = 0;
        if (rung == 0) nx = atomicAdd(w.counter, 1ull);
        constexpr unsigned kOne = kGroup >= 32 ? 0xffffffffu : ((1u << (kGroup & 31)) - 1u);
        item = n_groups + (long long)__shfl_sync(kOne << gs.gbase_lane, nx, gs.gbase_lane);  // the group finished together
      }
      first = false;
      if (!take()) return;           // no chain left
      if (state == ST_INIT) return;  // bound to a chain with candidates
    }
  };
  next_chain(true);
  for (;;) {
    if (__all_sync(kFull, state == ST_DONE)) break;
    if (state == ST_INIT) {
      jsel = w.sel_idx[(((size_t)ep * kOuter + w.it) * w.set.L + lrel) * p.K + rank];
#pragma unroll
      for (int i = 0; i < 4; ++i) c.xt[i] = R(lap[(size_t)i * w.set.Tcap + jsel]);
      gs.start(p);
      state = ST_RUN;
    }
    if (gs.trip(p)) {
      const size_t cslot = ((size_t)ep * w.set.L + lrel) * p.K + rank;
      const size_t tslot = (((size_t)ep * kOuter + w.it) * w.set.L + lrel) * p.K + rank;
      gs.store(w.cand_x + cslot * 4 * (N + 1), w.cand_u + cslot * 2 * N);
      if (rung == 0) {
        my_iters += (unsigned long long)c.iters;
        my_solves += 1;
        R xN[4];
        gs.terminal_state(xN);
        w.cand_cost[tslot] = relax_cost(p, xN, c.xt, (double)(T - 1 - jsel), w.scale);
        w.cand_lamb[tslot] = (double)c.lamb;
        w.cand_iters[tslot] = c.iters;
      }
      ++rank;
      if (rank < cnt) {
        state = ST_INIT;  // lambda is carried to the lap's next candidate
      } else {
        next_chain(false);
      }
    }
    __syncwarp();  // the finished trajectory is read above before the next candidate's start overwrites trial buffer 0
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_iters += __shfl_xor_sync(kFull, my_iters, o);
    my_solves += __shfl_xor_sync(kFull, my_solves, o);
  }
  if (lane == 0 && w.stats != nullptr && (my_solves | my_iters)) {
    atomicAdd(w.stats + 0, my_iters);
    atomicAdd(w.stats + 1, my_solves);
  }
}

// ------------------------------------------------------------------------------------------------ K1
// One warp per (episode, active lap).  The lap's SoA rows are read coalesced, the 1-norm distances
// ((|dx|+|dy|)+|dv|)+|dtheta| (same association as scipy's norm, bit-exact) are staged in shared
// memory, and K (+1) rounds of a warp-shuffle arg-min on the key (distance, index) emit the indices in rank
// order -- the order is semantically significant (lambda-chain order, winner rule).
//
// Ties.  The reference sorts with numpy.argsort's default kind (utils/base.py:339), an introsort that leaves EQUAL
// distances in an order which is a deterministic function of the whole array -- and the stock initial lap ends in
// three identical rows, so ties do reach the selection.  When no two of the K+1 smallest distances are equal the
// first K of any sort agree and the shuffle rounds are the answer; when a tie touches them, lane 0 replays NumPy's
// scalar algorithm (np_argsort_device: median-of-3 quicksort above 16 elements, insertion sort below, heapsort past
// the depth budget -- the published npysort `aquicksort`, checked against NumPy itself through the oracle's copy) on
// the staged distances, so the emitted indices are NumPy's in that case too.  A rare, sequential path: ~T log T
// shared-memory operations by one lane.
__device__ __forceinline__ bool np_less(double a, double b) { return a < b || (b != b && a == a); }

__device__ __noinline__ void np_aheapsort_device(const double* v, int* tosort, int n) {
  int* a = tosort - 1;  // 1-based
  int i, j, l, tmp;
  for (l = n >> 1; l > 0; --l) {
    tmp = a[l];
    for (i = l, j = l << 1; j <= n;) {
      if (j < n && np_less(v[a[j]], v[a[j + 1]])) j += 1;
      if (np_less(v[tmp], v[a[j]])) {
        a[i] = a[j];
        i = j;
        j += j;
      } else {
        break;
      }
    }
    a[i] = tmp;
  }
  for (; n > 1;) {
    tmp = a[n];
    a[n] = a[1];
    n -= 1;
    for (i = 1, j = 2; j <= n;) {
      if (j < n && np_less(v[a[j]], v[a[j + 1]])) j++;
      if (np_less(v[tmp], v[a[j]])) {
        a[i] = a[j];
        i = j;
        j += j;
      } else {
        break;
      }
    }
    a[i] = tmp;
  }
}

__device__ __noinline__ void np_argsort_device(const double* v, int num, int* t) {
  for (int j = 0; j < num; ++j) t[j] = j;
  if (num < 2) return;
  constexpr int kSmall = 15;  // ranges of up to 16 elements go to the insertion sort
  int stack[64], depth[32];
  int sp = 0, dp = 0, cdepth = 0;
  for (int n = num; n > 1; n >>= 1) cdepth += 2;  // 2 * floor(log2 num)
  int pl = 0, pr = num - 1;
  for (;;) {
    if (cdepth < 0) {
      np_aheapsort_device(v, t + pl, pr - pl + 1);
    } else {
      while (pr - pl > kSmall) {
        const int pm = pl + ((pr - pl) >> 1);
        int tmp;
        if (np_less(v[t[pm]], v[t[pl]])) { tmp = t[pm]; t[pm] = t[pl]; t[pl] = tmp; }
        if (np_less(v[t[pr]], v[t[pm]])) { tmp = t[pr]; t[pr] = t[pm]; t[pm] = tmp; }
        if (np_less(v[t[pm]], v[t[pl]])) { tmp = t[pm]; t[pm] = t[pl]; t[pl] = tmp; }
        const double vp = v[t[pm]];
        int pi = pl, pj = pr - 1;
        tmp = t[pm]; t[pm] = t[pj]; t[pj] = tmp;
        for (;;) {
          do ++pi;
          while (np_less(v[t[pi]], vp));
          do --pj;
          while (np_less(vp, v[t[pj]]));
          if (pi >= pj) break;
          tmp = t[pi]; t[pi] = t[pj]; t[pj] = tmp;
        }
        tmp = t[pi]; t[pi] = t[pr - 1]; t[pr - 1] = tmp;
        if (pi - pl < pr - pi) {  // larger part on the stack
          stack[sp++] = pi + 1;
          stack[sp++] = pr;
          pr = pi - 1;
        } else {
          stack[sp++] = pl;
          stack[sp++] = pi - 1;
          pl = pi + 1;
        }
        depth[dp++] = --cdepth;
      }
      for (int pi = pl + 1; pi <= pr; ++pi) {
        const int vi = t[pi];
        const double vp = v[vi];
        int pj = pi;
        while (pj > pl && np_less(vp, v[t[pj - 1]])) {
          t[pj] = t[pj - 1];
          --pj;
        }
        t[pj] = vi;
      }
    }
    if (sp == 0) break;
    pr = stack[--sp];
    pl = stack[--sp];
    cdepth = depth[--dp];
  }
}

constexpr int kKnnCountLimit = 512;  // laps up to this many points are ranked by counting, longer ones by arg-min rounds

struct KnnArgs {
  SafeSet set;
  int K, it, n_episodes;
  const double* x;        // (E,4) current states              (outer iteration 0)
  const double* x_pred;   // (E,4,N+1) previous winner          (outer iterations 1,2: its last column)
  int N;
  const int* active;      // (E) or null
  const int* ep_list;     // compacted episode list or null
  const int* n_ep;        // device: entries of ep_list
  long long* sel_idx;     // (E,3,L,K)
  int* sel_cnt;           // (E,L)
  // single-query form (i2lqr_select_knn): episode / absolute lap / explicit x0
  int q_episode, q_lap;
  const double* q_x0;
  long long* q_idx;
  int* q_cnt;
};

// K-nearest selection of one (episode, lap slot) by one warp; `dist` holds Tcap doubles and `order` Tcap ints of
// (shared) scratch.  Writes the indices in rank order to out[0..K) (-1 beyond the lap's length) and the count.
__device__ __forceinline__ void knn_warp(const SafeSet& set, int K, int e, int slot, const double (&g)[4], long long* out, int* out_cnt,
                                         double* dist, int* order, int lane) {
  const int T = set.lap_len[(size_t)e * set.L + slot];
  const double* lap = set.lap(e, slot);
  const size_t Tc = set.Tcap;
  for (int j = lane; j < T; j += 32)
    dist[j] = ((fabs(lap[j] - g[0]) + fabs(lap[Tc + j] - g[1])) + fabs(lap[2 * Tc + j] - g[2])) +
              fabs(lap[3 * Tc + j] - g[3]);
  __syncwarp();
  const int n = K < T ? K : T;
  bool tie = false;
  if (T <= kKnnCountLimit) {
    // Short laps (every lap of the reference's scenarios): rank by counting.  Element j's rank in the order
    // (distance, index) is below = #{d' < d_j} plus the equal distances with a smaller index; it is emitted if
    // rank < K.  T^2 / 32 comparisons per lane on shared memory, no dependent shuffle rounds -- at K = 64 the K + 1
    // arg-min rounds below are ~20 us of a 2.5 ms controller step, this is ~2 us.  A tie touches the K + 1 smallest
    // exactly when some distance occurs more than once and fewer than K distances lie below it.
    for (int j = lane; j < T; j += 32) {
      const double d = dist[j];
      int below = 0, equal_before = 0, equal_all = 0;
      for (int i = 0; i < T; ++i) {
        const double di = dist[i];
        below += di < d;
        equal_all += di == d;
        equal_before += (di == d) & (i < j);
      }
      const int rank = below + equal_before;
      if (d == d && rank < n) out[rank] = j;  // (a NaN distance never ranks; the arg-min rounds below skip it as well)
      tie = tie || (equal_all > 1 && below < K && below + 1 < T);
    }
    tie = __any_sync(0xffffffffu, tie);
  } else {
    double dprev = -INFINITY;
    int jprev = -1;
    const int rounds = n < T ? n + 1 : n;  // one more than emitted: does the K-th distance tie with the next?
    for (int r = 0; r < rounds; ++r) {
      double dbest = INFINITY;
      int jbest = 0x7fffffff;
      for (int j = lane; j < T; j += 32) {
        const double d = dist[j];
        const bool after = d > dprev || (d == dprev && j > jprev);
        if (after && (d < dbest || (d == dbest && j < jbest))) {
          dbest = d;
          jbest = j;
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double d2 = __shfl_xor_sync(0xffffffffu, dbest, o);
        const int j2 = __shfl_xor_sync(0xffffffffu, jbest, o);
        if (d2 < dbest || (d2 == dbest && j2 < jbest)) {
          dbest = d2;
          jbest = j2;
        }
      }
      tie = tie || (r > 0 && dbest == dprev);
      if (lane == 0 && r < n) out[r] = jbest;
      dprev = dbest;
      jprev = jbest;
    }
  }
  if (tie) {  // warp-uniform in both forms
    __syncwarp();  // the indices written above by other lanes are overwritten below by lane 0
    if (lane == 0) {
      np_argsort_device(dist, T, order);
      for (int r = 0; r < n; ++r) out[r] = order[r];
    }
    __syncwarp();
  }
  if (lane == 0) {
    *out_cnt = n;
    for (int r = n; r < K; ++r) out[r] = -1;
  }
}

__global__ void knn_kernel(const __grid_constant__ KnnArgs a) {
  extern __shared__ double dist_smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  // per warp: Tcap distances, then Tcap ints for the tie path's index array
  double* dist = dist_smem + (size_t)wib * (a.set.Tcap + (a.set.Tcap + 1) / 2);
  int* order = reinterpret_cast<int*>(dist + a.set.Tcap);
  const long long wid = (long long)blockIdx.x * warps_per_block + wib;
  int e, lrel, slot;
  double g[4];
  long long* out;
  int* out_cnt;
  if (a.q_x0 != nullptr) {
    if (wid != 0) return;
    e = a.q_episode;
    slot = a.q_lap % a.set.L;
    lrel = 0;
    for (int i = 0; i < 4; ++i) g[i] = a.q_x0[i];
    out = a.q_idx;
    out_cnt = a.q_cnt;
  } else {
    if (wid >= (long long)(a.n_ep != nullptr ? *a.n_ep : a.n_episodes) * a.set.L) return;
    e = (int)(wid / a.set.L);
    if (a.ep_list != nullptr) e = a.ep_list[e];
    lrel = (int)(wid % a.set.L);
    const int count = a.set.lap_count[e];
    const int n_act = count < a.set.L ? count : a.set.L;
    out_cnt = a.sel_cnt + (size_t)e * a.set.L + lrel;
    if (lrel >= n_act || (a.active != nullptr && !a.active[e])) {
      if (lane == 0) *out_cnt = 0;
      return;
    }
    const int min_iter = count - a.set.L > 0 ? count - a.set.L : 0;
    slot = (min_iter + lrel) % a.set.L;
    for (int i = 0; i < 4; ++i)  // x_guess: utils/base.py:398-401
      g[i] = a.it == 0 ? a.x[(size_t)e * 4 + i] : a.x_pred[((size_t)e * 4 + i) * (a.N + 1) + a.N];
    out = a.sel_idx + (((size_t)e * kOuter + a.it) * a.set.L + lrel) * a.K;
  }
  knn_warp(a.set, a.K, e, slot, g, out, out_cnt, dist, order, lane);
}

// ------------------------------------------------------------------------------------------------ K3
// One thread per episode: `min(cost_list)` is Python's lexicographic list comparison over the per-lap cost
// vectors (first differing element decides, a prefix is smaller), then the first arg-min inside the
// winning lap (utils/base.py:462-465); the winner's trajectory becomes x_pred / u_pred.
struct WinnerArgs {
  int n_episodes, it, L, K, N;
  const int* lap_count;
  const int* lap_len;      // (E,L)
  const int* active;
  const int* ep_list;      // compacted episode list or null
  const int* n_ep;         // device: entries of ep_list
  const int* sel_cnt;      // (E,L)
  const long long* sel_idx;
  const double* cand_cost; // (E,3,L,K)
  const double* cand_x;
  const double* cand_u;
  double* u;               // (E,2)
  double* x_pred;          // (E,4,N+1)
  double* u_pred;          // (E,2,N)
  int* winner_lap;
  int* winner_rank;
  long long* winner_idx;
  double* winner_cost;
  int* shrink;
  int stage_costs;         // 1: dynamic shared memory holds (warps per block) x L x K doubles
};

// Winner of episode e in outer iteration a.it, by one warp; `stage` = L*K doubles of shared scratch or null.
__device__ __forceinline__ void winner_warp(const WinnerArgs& a, int e, int lane, double* stage) {
  const int count = a.lap_count[e];
  const int n_act = count < a.L ? count : a.L;
  const int min_iter = count - a.L > 0 ? count - a.L : 0;
  int bl = 0, br = 0;
  // The rule below is a chain of dependent reads (a comparison stops at the first differing element), ~1 us each from
  // global memory; when the step's cost table fits, all lanes first copy it to shared memory in one parallel sweep.
  const double* cost = a.cand_cost + ((size_t)e * kOuter + a.it) * a.L * a.K;
  if (stage != nullptr) {
    for (int i = lane; i < n_act * a.K; i += 32) stage[i] = cost[i];
    __syncwarp();
    cost = stage;
  }
  if (lane == 0) {
    const int* cnt = a.sel_cnt + (size_t)e * a.L;
    for (int l = 1; l < n_act; ++l) {
      const double* ca = cost + (size_t)l * a.K;
      const double* cb = cost + (size_t)bl * a.K;
      const int na = cnt[l], nb = cnt[bl];
      const int n = na < nb ? na : nb;
      bool less = na < nb;
      for (int i = 0; i < n; ++i)
        if (ca[i] != cb[i]) {
          less = ca[i] < cb[i];
          break;
        }
      if (less) bl = l;
    }
    const double* cv = cost + (size_t)bl * a.K;
    for (int r = 1; r < cnt[bl]; ++r)
      if (cv[r] < cv[br]) br = r;
    const long long j = a.sel_idx[(((size_t)e * kOuter + a.it) * a.L + bl) * a.K + br];
    a.winner_lap[e] = min_iter + bl;
    a.winner_rank[e] = br;
    a.winner_idx[e] = j;
    a.winner_cost[e] = cv[br];
    const int T = a.lap_len[(size_t)e * a.L + (min_iter + bl) % a.L];
    a.shrink[e] = (j + 1) > (long long)(T - 1);  // utils/base.py:474-477
  }
  bl = __shfl_sync(0xffffffffu, bl, 0);
  br = __shfl_sync(0xffffffffu, br, 0);
  const size_t cslot = ((size_t)e * a.L + bl) * a.K + br;
  const int nx = 4 * (a.N + 1), nu = 2 * a.N;
  for (int i = lane; i < nx; i += 32) a.x_pred[(size_t)e * nx + i] = a.cand_x[cslot * nx + i];
  for (int i = lane; i < nu; i += 32) a.u_pred[(size_t)e * nu + i] = a.cand_u[cslot * nu + i];
  if (lane < 2) a.u[(size_t)e * 2 + lane] = a.cand_u[cslot * nu + (size_t)lane * a.N];
}

__global__ void winner_kernel(const __grid_constant__ WinnerArgs a) {
  // one warp per episode: lane 0 decides, all lanes copy
  int e = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= (a.n_ep != nullptr ? *a.n_ep : a.n_episodes)) return;
  if (a.ep_list != nullptr) e = a.ep_list[e];
  if (a.active != nullptr && !a.active[e]) return;
  extern __shared__ double winner_smem[];
  winner_warp(a, e, lane, a.stage_costs ? winner_smem + (size_t)(threadIdx.x >> 5) * a.L * a.K : nullptr);
}

// ------------------------------------------------------------------------------- split controller step
// One controller step whose laps are solved by several handles (GPUs): SURVEY.md 8e "Collective".  Handle
// `part` of `n_parts` owns the active laps with (relative lap number) % n_parts == part and runs their lambda
// chains; what the other handles need of a lap travels as one packed row per (episode, lap):
//   [0] rank of the lap's first arg-min candidate   [1..K] relaxed costs   [K+1..2K] lambda after each candidate
//   [2K+1..3K] executed iterations   then that candidate's xvar (4,N+1) and uvar (2,N)
// -- only a lap's first arg-min can win the step (utils/base.py:462-465), so its trajectory is the only one that
// has to cross.  After the all-gather every handle holds every row, unpacks them into its own step workspace
// and runs the same winner_kernel: all handles end the outer iteration with bit-identical x_pred / u_pred.
struct SplitArgs {
  int n_episodes, it, L, K, N;
  int part, n_parts;
  const int* lap_count;    // (E)
  const int* sel_cnt;      // (E,L)  true K-nearest counts
  int* sel_cnt_own;        // (E,L)  counts with the laps of other handles zeroed (what the chains walk)
  double* cand_cost;       // (E,3,L,K)
  double* cand_lamb;
  int* cand_iters;
  double* cand_x;          // (E,L,K,4,N+1)
  double* cand_u;          // (E,L,K,2,N)
  double* rows;            // pack: (E,L,row) of this handle; unpack: (n_parts,E,L,row) gathered
};
__host__ __device__ __forceinline__ int split_row_doubles(int K, int N) { return 1 + 3 * K + 4 * (N + 1) + 2 * N; }

__global__ void split_mask_kernel(const __grid_constant__ SplitArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_episodes * a.L) return;
  a.sel_cnt_own[i] = (i % a.L) % a.n_parts == a.part ? a.sel_cnt[i] : 0;
}

// one warp per (episode, lap)
__global__ void split_pack_kernel(const __grid_constant__ SplitArgs a) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)a.n_episodes * a.L) return;
  const int e = (int)(wid / a.L), l = (int)(wid % a.L);
  const int row = split_row_doubles(a.K, a.N);
  double* out = a.rows + (size_t)wid * row;
  const int cnt = l % a.n_parts == a.part ? a.sel_cnt[wid] : 0;
  if (cnt == 0) {
    for (int i = lane; i < row; i += 32) out[i] = 0.0;
    return;
  }
  const size_t tslot = (((size_t)e * kOuter + a.it) * a.L + l) * a.K;
  int br = 0;
  for (int r = 1; r < cnt; ++r)  // first arg-min, the comparison winner_kernel makes
    if (a.cand_cost[tslot + r] < a.cand_cost[tslot + br]) br = r;
  if (lane == 0) out[0] = (double)br;
  for (int r = lane; r < a.K; r += 32) {
    out[1 + r] = a.cand_cost[tslot + r];
    out[1 + a.K + r] = a.cand_lamb[tslot + r];
    out[1 + 2 * a.K + r] = (double)a.cand_iters[tslot + r];
  }
  const size_t cslot = ((size_t)e * a.L + l) * a.K + br;
  const int nx = 4 * (a.N + 1), nu = 2 * a.N;
  for (int i = lane; i < nx; i += 32) out[1 + 3 * a.K + i] = a.cand_x[cslot * nx + i];
  for (int i = lane; i < nu; i += 32) out[1 + 3 * a.K + nx + i] = a.cand_u[cslot * nu + i];
}

__global__ void split_unpack_kernel(const __grid_constant__ SplitArgs a) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)a.n_episodes * a.L) return;
  const int e = (int)(wid / a.L), l = (int)(wid % a.L);
  const int owner = l % a.n_parts;
  if (owner == a.part || a.sel_cnt[wid] == 0) return;  // own rows are already in place
  const int row = split_row_doubles(a.K, a.N);
  const double* in = a.rows + ((size_t)owner * a.n_episodes * a.L + wid) * row;
  const size_t tslot = (((size_t)e * kOuter + a.it) * a.L + l) * a.K;
  const int br = (int)in[0];
  for (int r = lane; r < a.K; r += 32) {
    a.cand_cost[tslot + r] = in[1 + r];
    a.cand_lamb[tslot + r] = in[1 + a.K + r];
    a.cand_iters[tslot + r] = (int)in[1 + 2 * a.K + r];
  }
  const size_t cslot = ((size_t)e * a.L + l) * a.K + br;
  const int nx = 4 * (a.N + 1), nu = 2 * a.N;
  for (int i = lane; i < nx; i += 32) a.cand_x[cslot * nx + i] = in[1 + 3 * a.K + i];
  for (int i = lane; i < nu; i += 32) a.cand_u[cslot * nu + i] = in[1 + 3 * a.K + nx + i];
}

// ------------------------------------------------------------------------------------------------ K4
// Batched closed-loop episodes: the plant, obstacle motion, termination test, playback bookkeeping and
// the lap append all stay on the device (Simulator.sim utils/base.py:708-726, KineticBicycle.forward_one_step
// :157-213, Obstacle.update_obstacle / reset_obstacle :50-64, calc_input's playback branch :377-382,
// iterative_ilqr/tests/ilqr_test.py:56-60).
struct EpisodeState {
  double* x;            // (E,4) plant state == controller state
  double* upred;        // (E,2,N) u_pred of the last solve step (playback source)
  int* step;            // steps taken in the current lap
  int* done;            // lap finished
  int* horizon;         // controller num_horizon (N when solving, < N while playing back)
  int* pb;              // next column of upred to play back
  int* active;          // solves this step
  int* status;          // 0 ok, 1 lap hit max_sim_steps, 2 playback ran out of inputs, 3 lap too short / long to store
  double* lap_x;        // (E,S+1,4)
  double* lap_u;        // (E,S,2)
  int S;
  DevObstacle* obs;     // current obstacle per episode (moves)
  const DevObstacle* obs0;
  DevObstacle* obs_extra;         // (E, kMaxExtraObstacles) or null
  const DevObstacle* obs_extra0;
  const int* n_extra;
  const double* goal;   // (4)
  int* n_running;       // episodes whose lap is still running
};

// Also compacts the episodes that solve this step into `ep_list` (warp-aggregated append; the order is arbitrary and
// does not matter, every episode's step is computed independently of where it sits in the list) and publishes the
// counts to mapped host memory, which the host reads WITHOUT synchronising: both counts only fall during a lap, so
// a stale value is an upper bound -- good enough to size the next launches and to stop issuing steps.
__global__ void episode_pre_kernel(EpisodeState s, int E, int N_param, int* ep_list, int* n_ep) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = e < E && !s.done[e] && s.horizon[e] == N_param;
  if (e < E) s.active[e] = act ? 1 : 0;
  const unsigned m = __ballot_sync(0xffffffffu, act);
  if (m == 0) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(n_ep, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (act) ep_list[base + __popc(m & ((1u << lane) - 1u))] = e;
}
__global__ void episode_publish_kernel(const int* n_running, const int* n_ep, volatile int* host_words) {
  host_words[0] = *n_running;
  host_words[1] = *n_ep;
  __threadfence_system();
}

// Philox4x32-10 counter-based generator (Salmon et al. 2011): one 128-bit block per (episode, lap, step).
__device__ __forceinline__ void philox4x32(uint32_t key0, uint32_t key1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ key0, n2 = hi0 ^ c3 ^ key1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    key0 += 0x9E3779B9u;
    key1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

__global__ void episode_post_kernel(EpisodeState s, const __grid_constant__ KParams p, const double* u_solve,
                                    const double* upred_solve, const int* shrink, int E, int sim_i, double goal_tol,
                                    unsigned long long noise_seed, int lap_index) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E || s.done[e]) return;
  const int N = p.N;
  double ua, ud;
  if (s.active[e]) {
    ua = u_solve[(size_t)e * 2];
    ud = u_solve[(size_t)e * 2 + 1];
    for (int i = 0; i < 2 * N; ++i) s.upred[(size_t)e * 2 * N + i] = upred_solve[(size_t)e * 2 * N + i];
    s.pb[e] = 1;
    if (shrink[e]) s.horizon[e] = N - 1;  // utils/base.py:474-477
  } else {
    const int c = s.pb[e];
    if (c >= N) {  // the reference would index an empty u_old here (IndexError)
      s.status[e] = 2;
      s.done[e] = 1;
      atomicSub(s.n_running, 1);
      return;
    }
    ua = s.upred[(size_t)e * 2 * N + c];      // utils/base.py:378-381
    ud = s.upred[(size_t)e * 2 * N + N + c];
    s.pb[e] = c + 1;
    s.horizon[e] -= 1;
  }
  // plant: kinetic_bicycle (systems/kinetic_bicycle.py:10-27), noise off
  double* x = s.x + (size_t)e * 4;
  double sn, cs;
  sincos_fast(x[3], &sn, &cs);
  const double sd = x[2] * p.dt + (ua * p.dt2) / 2.0;
  const double x0 = x[0] + cs * sd, x1 = x[1] + sn * sd;
  double x2 = x[2] + ua * p.dt, x3 = x[3] + ud * p.dt;
  if (noise_seed != 0) {
    // plant noise of KineticBicycle.forward_dynamics (utils/base.py:207-211): clipped Gaussians on speed and
    // heading; two normals per step by Box-Muller from one Philox block keyed by (seed, episode, lap, step)
    uint32_t r[4];
    philox4x32((uint32_t)noise_seed, (uint32_t)(noise_seed >> 32), (uint32_t)e, (uint32_t)lap_index, (uint32_t)s.step[e], 0x1234567u, r);
    const double u1 = ((double)r[0] + 1.0) * (1.0 / 4294967296.0), u2 = (double)r[1] * (1.0 / 4294967296.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn2, cs2;
    sincos(6.283185307179586 * u2, &sn2, &cs2);
    const double noise_v = fmax(-0.05, fmin(rad * cs2 * 0.01, 0.05));
    const double noise_theta = fmax(-0.05, fmin(rad * sn2 * 0.005, 0.05));
    x2 += 0.5 * noise_v;
    x3 += 0.5 * noise_theta;
  }
  x[0] = x0;
  x[1] = x1;
  x[2] = x2;
  x[3] = x3;
  const int st = s.step[e];
  double* lx = s.lap_x + ((size_t)e * (s.S + 1) + st + 1) * 4;
  lx[0] = x0;
  lx[1] = x1;
  lx[2] = x2;
  lx[3] = x3;
  s.lap_u[((size_t)e * s.S + st) * 2] = ua;
  s.lap_u[((size_t)e * s.S + st) * 2 + 1] = ud;
  s.step[e] = st + 1;
  DevObstacle* o = s.obs + e;
  if (o->present) {  // Obstacle.update_obstacle
    if (o->moving_option == 1) o->y += o->spd * p.dt;
    if (o->moving_option == 2) o->x -= o->spd * p.dt;
  }
  if (s.n_extra != nullptr) {
    for (int m = 0; m < s.n_extra[e]; ++m) {
      DevObstacle* q = s.obs_extra + (size_t)e * kMaxExtraObstacles + m;
      if (!q->present) continue;
      if (q->moving_option == 1) q->y += q->spd * p.dt;
      if (q->moving_option == 2) q->x -= q->spd * p.dt;
    }
  }
  const double d0 = x0 - s.goal[0], d1 = x1 - s.goal[1], d2 = x2 - s.goal[2], d3 = x3 - s.goal[3];
  const bool reached = sqrt(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) <= goal_tol;
  if (reached || sim_i == s.S - 1) {
    if (!reached) s.status[e] = 1;
    s.done[e] = 1;
    atomicSub(s.n_running, 1);
  }
}

// One warp per episode: snap the lap's last row to the goal, append it to the safe-set ring, reset the
// episode for its next lap (state <- 0 as update_memory_post_iter does, utils/base.py:195).
__global__ void episode_finish_kernel(EpisodeState s, SafeSet set, int E, int N_param, int* lap_len_out) {
  const int e = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  const int steps = s.step[e];
  const int T = steps + 1;
  double* lx = s.lap_x + (size_t)e * (s.S + 1) * 4;
  if (lane < 4) lx[(size_t)steps * 4 + lane] = s.goal[lane];
  __syncwarp();
  const bool storable = T > N_param && T <= set.Tcap;
  if (storable) {
    const int count = set.lap_count[e];
    const int slot = count % set.L;
    double* dst = set.ss + ((size_t)e * set.L + slot) * 4 * (size_t)set.Tcap;
    for (int j = lane; j < T; j += 32)
      for (int i = 0; i < 4; ++i) dst[(size_t)i * set.Tcap + j] = lx[(size_t)j * 4 + i];
    __syncwarp();
    if (lane == 0) {
      set.lap_len[(size_t)e * set.L + slot] = T;
      set.lap_count[e] = count + 1;
    }
  }
  if (lane == 0) {
    if (!storable) s.status[e] = 3;
    lap_len_out[e] = steps;
    s.step[e] = 0;
    s.done[e] = 0;
    s.horizon[e] = N_param;
    s.pb[e] = 0;
    s.obs[e] = s.obs0[e];
    if (s.n_extra != nullptr)
      for (int m = 0; m < kMaxExtraObstacles; ++m) s.obs_extra[(size_t)e * kMaxExtraObstacles + m] = s.obs_extra0[(size_t)e * kMaxExtraObstacles + m];
    for (int i = 0; i < 4; ++i) s.x[(size_t)e * 4 + i] = 0.0;
  }
}

// ---------------------------------------------------------------------------------- FP64 FMA-pipe peak
// Dependent-free DFMA streams: the measured denominator of the roofline (MEASURED_PEAKS.json carries no
// FP64 figure).  8 independent accumulators per thread, `iters` rounds of 8 FMAs.
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
  for (int i = 0; i < iters; ++i) {
    r0 = fma(r0, a, b);
    r1 = fma(r1, a, b);
    r2 = fma(r2, a, b);
    r3 = fma(r3, a, b);
    r4 = fma(r4, a, b);
    r5 = fma(r5, a, b);
    r6 = fma(r6, a, b);
    r7 = fma(r7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
}

}  // namespace i2lqr
