// spec_step.cuh -- one controller step (iLqr.calc_input, utils/base.py:371-479) as ONE persistent launch, with the
// candidates of a lambda chain solved speculatively on the SMs the chains themselves leave idle.
//
// The K candidates of one (episode, lap) are a sequential chain only through lambda: candidate j+1 starts from the
// lambda candidate j ends with (utils/base.py:393,414-426), and a candidate's result is a pure function of
// (candidate, lambda_in).  lambda moves by /10 and *10 from 1, so in a run it takes few distinct doubles: over the
// recorded reference runs (tests/golden/steps_*.npz, 43 186 candidate solves) the 16 most frequent bit patterns
// cover 91 % of all lambda_in, the top 24 cover 96 %.  A controller step has 8 chains and a B200 has 592 SM
// sub-partitions, so while a chain's warp (the "resolver") walks its candidates in order, "worker" warps solve
// every candidate j >= 1 of every chain for each of the S most likely lambda_in values -- the speculation set, the
// exact doubles seen most often in earlier steps (a 64-entry histogram kept on the device).  When the resolver
// reaches candidate j it looks for an entry whose lambda_in equals the chain's lambda BIT FOR BIT: done -> take
// its result; in progress -> wait for it; not started -> claim it and solve it itself; not in the set (a miss) ->
// solve it itself.  Every result is produced by the same GroupSolver code from the same operands, so the step is
// bit-identical to the sequential chain whatever hits or misses; speculation only changes when a result exists.
//
// The step's three outer iterations (K1 K-nearest -> K2 chains -> K3 winner) run inside the same launch: resolver
// warps do their lap's K-nearest selection, the last resolver of an episode to finish an outer iteration applies the
// winner rule and releases the others.  No cooperative launch is needed: roles are taken from a ticket counter by
// the blocks that actually run, workers never block (they poll), and a resolver only ever waits for an entry that
// a running worker has claimed, or for resolvers that hold tickets (hence run).
#pragma once
#include "kernels.cuh"

namespace i2lqr {

constexpr int kSpecMaxSet = 32;      // speculation set entries (one per lane of the resolver)
constexpr int kSpecHist = 64;        // histogram slots (two per lane)
constexpr int kSpecRungs = 4;
#ifndef I2LQR_SPEC_WG
#define I2LQR_SPEC_WG 4
#endif
constexpr int kSpecWorkerGroup = I2LQR_SPEC_WG;  // lanes per speculative candidate: 8 candidates per worker warp (compile-time knob for A-B builds)
constexpr int kSpecPubWords = 6;
constexpr int kSpecInlineEpisodes = 8;  // steps of up to this many episodes need no copy at all: state in the launch parameters, results by zero-copy stores
// per episode in host_out: u (2), x_pred (4 (N+1)), u_pred (2 N), winner cost | winner index | lap, rank | shrink
__host__ __device__ __forceinline__ int spec_host_words(int N) { return 2 + 4 * (N + 1) + 2 * N + 1 + 1 + 1 + 1; }
constexpr unsigned long long kSpecUnused = 0x7ff8dead00000000ull;  // NaN payload: never a lambda

struct SpecArgs {
  // ---- scope
  int n_episodes;              // episodes (host upper bound when ep_list is used)
  const int* ep_list;          // compacted episode list or null
  const int* n_ep;             // device: entries of ep_list
  SafeSet set;
  const double* x;             // (E,4)
  int n_inline;                // > 0: the states of episodes 0 .. n_inline-1 travel in the launch parameters (x_inline) instead
  double x_inline[4 * kSpecInlineEpisodes];
  unsigned long long* host_out;  // mapped pinned host memory or null: results of a small step written straight to the host
                                 // ([0] flag = epoch, written last; [1] iterations [2] solves; then per episode kSpecHostWords(N) words)
  const DevObstacle* ep_obs;   // (E)
  const DevObstacle* ep_obs_extra;  // (E, kMaxExtraObstacles) or null
  const int* ep_n_extra;       // (E)
  int S;                       // speculation set size in use, <= kSpecMaxSet
  int band;                    // the workers start the `band` most frequent values of every (rank, chain) before the rest (0 / >= S: one band)
  int c_max;                   // chains the tables are laid out for (>= n_episodes * L)
  int e_stride;                // episodes the per-episode control words are laid out for (>= n_episodes)
  // ---- step workspace (the arrays the three-kernel form uses)
  long long* sel_idx;          // (E,3,L,K)
  int* sel_cnt;                // (E,L)
  double* cand_x;              // (E,L,K,4,N+1): only each lap's running best is stored
  double* cand_u;              // (E,L,K,2,N)
  double* cand_cost;           // (E,3,L,K)
  double* cand_lamb;
  int* cand_iters;
  WinnerArgs win;              // it / ep_list fields are filled in by the kernel
  // ---- speculation tables: entry ((it * c_max + chain) * K + j) * S + s
  unsigned* item_state;        // (epoch << 2) | 1 claimed / 2 done; anything below (epoch << 2) is free.  Values only grow (the
                               // host clears the table when the epoch wraps), so a claim is ONE atomicMax with no prior read
  double* item_lamb;           // lambda_out
  double* item_cost;           // relaxed cost
  int* item_iters;
  double* item_traj;           // xvar (4,N+1) then uvar (2,N)
  unsigned epoch;              // launch number, stamps every flag: nothing is cleared between launches
  // ---- control words (zero at allocation, restored by the last block to leave)
  unsigned* ctl;               // [0] ticket [1] blocks left [2] chains finished [3] all done (epoch) [4..6] queue heads
  unsigned long long* pub;     // (3, c_max, K, 6) what a worker needs to start candidate j, published by the chain's resolver after its
                               //   K-nearest: [0] (epoch << 32) | safe-set index (0xffffffff: the lap has no candidate j), written LAST
                               //   (release) so that it doubles as the "record is there" flag; [1..4] terminal point; [5] Qfun
  int k_eff;                   // host's upper bound of candidates per chain (min(K, longest active lap)): bounds the work list
  unsigned long long* walked;  // (3, c_max)  (epoch << 32) | candidates the resolver has passed
  unsigned* arrived;           // (3, E)      resolvers of the episode that finished the outer iteration
  unsigned* iter_done;         // (3, E)      epoch once the episode's winner of that outer iteration is published
  // ---- lambda_in histogram and the speculation set derived from it
  unsigned long long* hist_key;  // (kSpecHist) bit patterns
  unsigned* hist_cnt;            // (kSpecHist)
  unsigned long long* spec_set;  // (kSpecMaxSet) the S most frequent patterns, frozen for the launch
  unsigned long long* item_dbg;    // optional (developer): per table entry [start ns, end ns, rounds, iterations]
  unsigned long long* worker_dbg;  // optional (developer): per block [rounds with a live candidate, cycles picking entries, cycles in rounds, cycles publishing]
  unsigned long long* timeline;    // optional (developer): per (it, chain) 8 words of globaltimer ns / counters, see i2lqr_spec_timeline
  unsigned long long* step_stats;  // [0] += iLQR iterations, [1] += candidate solves of the step (the reference's path only)
  unsigned long long* stats;       // cumulative: [0] hits [1] waited for [2] self-solved (in the set, not started) [3] misses
                                   // [4] speculative solves finished [5] abandoned [6] launches
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// polls of flags that guard no data (step over? chain past this candidate?): relaxed, no L1 invalidation
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ double spec_x(const SpecArgs& a, int ep, int i) {
  return a.n_inline > 0 ? a.x_inline[ep * 4 + i] : a.x[(size_t)ep * 4 + i];
}
// Binds a solver's candidate to episode `ep` (state, obstacle); the terminal point and lambda follow per candidate.
template <class Solver>
__device__ __forceinline__ void spec_bind_episode(const KParams& p, const SpecArgs& a, int ep, Solver& gs) {
  load_obstacle(p, gs.c, a.ep_obs + ep);
  if (a.ep_n_extra != nullptr) {
    gs.c.ex = a.ep_obs_extra + (size_t)ep * kMaxExtraObstacles;
    gs.c.n_ex = a.ep_n_extra[ep];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) gs.c.x0[i] = spec_x(a, ep, i);
  sincos_fast(gs.c.x0[3], &gs.c.s0, &gs.c.c0);
}

struct SpecLap {  // the stored lap a chain draws its terminal points from
  const double* lap;
  int T;
};
__device__ __forceinline__ SpecLap spec_lap(const SpecArgs& a, int ep, int lrel) {
  const int count = a.set.lap_count[ep];
  const int min_iter = count - a.set.L > 0 ? count - a.set.L : 0;
  const int slot = (min_iter + lrel) % a.set.L;
  SpecLap l;
  l.T = a.set.lap_len[(size_t)ep * a.set.L + slot];
  l.lap = a.set.lap(ep, slot);
  return l;
}

#ifdef I2LQR_SPEC_MAXREG  // compile-time knob for A-B builds: registers per thread (more resident warps, i.e. lane groups)
#define I2_SPEC_BOUNDS __maxnreg__(I2LQR_SPEC_MAXREG)
#else
#define I2_SPEC_BOUNDS __launch_bounds__(kSolveThreads, 1)
#endif
template <bool kQR, int kWG = kSpecWorkerGroup>
__global__ void I2_SPEC_BOUNDS
spec_step_kernel(const __grid_constant__ KParams p, const __grid_constant__ SpecArgs a) {
  extern __shared__ double smem[];
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int S32 = kSolveThreads;
  const int lane = threadIdx.x & 31;
  const int N = p.N, K = p.K, L = a.set.L, S = a.S;
  const unsigned epoch = a.epoch;
  unsigned ticket = 0;
  if (lane == 0) ticket = atomicAdd(a.ctl + 0, 1u);
  ticket = __shfl_sync(kFull, ticket, 0);
  const int n_eps = a.n_ep != nullptr ? *a.n_ep : a.n_episodes;
  const int n_chains = n_eps * L;
  const unsigned long long spec_bits = lane < S ? a.spec_set[lane] : kSpecUnused;
  unsigned long long st_iters = 0, st_solves = 0, st_hit = 0, st_wait = 0, st_self = 0, st_miss = 0, st_spec = 0, st_drop = 0;
  const int n_traj = 6 * N + 4;

  if ((int)ticket < n_chains) {
    // =============================================================== resolver: owns chain `ticket`
    const int chain = (int)ticket, pos = chain / L, lrel = chain % L;
    const int ep = a.ep_list != nullptr ? a.ep_list[pos] : pos;
    GroupSolver<double, kQR, kSpecRungs, 32, S32> gs;
    gs.bind(smem, lane, N);
    spec_bind_episode(p, a, ep, gs);
    const int count = a.set.lap_count[ep];
    const int n_act = count < L ? count : L;
    const SpecLap lp = spec_lap(a, ep, lrel < n_act ? lrel : 0);
    unsigned long long hkey0 = a.hist_key[lane], hkey1 = a.hist_key[lane + 32];
    for (int it = 0; it < kOuter; ++it) {
      unsigned long long tl[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // start, knn done, rank 0 done, walk done, released | wait ns, in-place solves, their iterations
      if (a.timeline != nullptr) tl[0] = global_ns();
      const double scale = it == 0 ? 1.0 : (it == 1 ? 10.0 : 100.0);  // 10**it, utils/base.py:428-435
      const size_t trow = (((size_t)ep * kOuter + it) * L + lrel) * K;
      int cnt = 0;
      if (lrel < n_act) {  // K1 for this lap (utils/base.py:398-402)
        double g[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)  // (x_pred was written by another warp, possibly on another SM: read it from L2)
          g[i] = it == 0 ? spec_x(a, ep, i) : __ldcg(a.win.x_pred + ((size_t)ep * 4 + i) * (N + 1) + N);
        double* dist = smem;
        int* order = reinterpret_cast<int*>(dist + a.set.Tcap);
        const int count2 = a.set.lap_count[ep];
        const int min_iter = count2 - L > 0 ? count2 - L : 0;
        __syncwarp();
        knn_warp(a.set, K, ep, (min_iter + lrel) % L, g, a.sel_idx + trow, a.sel_cnt + (size_t)ep * L + lrel, dist, order, lane);
        __syncwarp();
        cnt = K < lp.T ? K : lp.T;
      } else {
        if (lane == 0) a.sel_cnt[(size_t)ep * L + lrel] = 0;
        for (int r = lane; r < K; r += 32) a.sel_idx[trow + r] = -1;
      }
      for (int r = cnt + lane; r < K; r += 32) {  // unused trace slots read as "no candidate" (all-ones patterns)
        a.cand_cost[trow + r] = __longlong_as_double(-1ll);
        a.cand_lamb[trow + r] = __longlong_as_double(-1ll);
        a.cand_iters[trow + r] = -1;
      }
      __threadfence();
      __syncwarp();
      unsigned long long* walked = a.walked + (size_t)it * a.c_max + chain;
      if (lane == 0) st_release_u64(walked, (unsigned long long)epoch << 32);
      {  // publish the candidates to the workers: terminal point and cost-to-go first, the stamped index word last
        unsigned long long* pub = a.pub + ((size_t)it * a.c_max + chain) * K * kSpecPubWords;
        for (int r = lane; r < cnt; r += 32) {
          const long long js = a.sel_idx[trow + r];
          double* rec = reinterpret_cast<double*>(pub + (size_t)r * kSpecPubWords);
#pragma unroll
          for (int i = 0; i < 4; ++i) rec[1 + i] = lp.lap[(size_t)i * a.set.Tcap + js];
          rec[5] = (double)(lp.T - 1 - js);
        }
        __threadfence();
        __syncwarp();
        for (int r = lane; r < K; r += 32)
          st_release_u64(pub + (size_t)r * kSpecPubWords,
                         ((unsigned long long)epoch << 32) | (r < cnt ? (unsigned long long)(unsigned)a.sel_idx[trow + r] : 0xffffffffull));
      }
      if (a.timeline != nullptr) tl[1] = global_ns();
      // ---- walk the chain in rank order, lambda carried (utils/base.py:393,403-455)
      double lamb = p.lamb0;
      int br = 0;
      double best_cost = 0.0;
      for (int j = 0; j < cnt; ++j) {
        if (lane == 0 && j > 0) st_release_u64(walked, ((unsigned long long)epoch << 32) | (unsigned)j);
        const long long jsel = a.sel_idx[trow + j];
        const double qfun = (double)(lp.T - 1 - jsel);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(lamb);
        const unsigned match = __ballot_sync(kFull, spec_bits == bits);
        // ---- histogram of lambda_in (next launches' speculation set); never on the result path
        if (j > 0) {
          const bool m0 = hkey0 == bits, m1 = hkey1 == bits;
          if (m0) atomicAdd(a.hist_cnt + lane, 1u);
          if (m1) atomicAdd(a.hist_cnt + lane + 32, 1u);
          if (__ballot_sync(kFull, m0 || m1) == 0) {  // new pattern: replace the least counted slot
            unsigned c0 = a.hist_cnt[lane], c1 = a.hist_cnt[lane + 32];
            unsigned best = c0 <= c1 ? c0 : c1;
            int slot = c0 <= c1 ? lane : lane + 32;
            for (int o = 16; o > 0; o >>= 1) {
              const unsigned ob = __shfl_xor_sync(kFull, best, o);
              const int os = __shfl_xor_sync(kFull, slot, o);
              if (ob < best || (ob == best && os < slot)) {
                best = ob;
                slot = os;
              }
            }
            if (lane == 0) {
              const unsigned long long old = a.hist_key[slot];
              if (atomicCAS(a.hist_key + slot, old, bits) == old) a.hist_cnt[slot] = 1u;
            }
            __syncwarp();
            hkey0 = a.hist_key[lane];
            hkey1 = a.hist_key[lane + 32];
          }
        }
        int how = 0;  // 0 solve here, 1 take the table entry
        size_t idx = 0;
        if (j > 0 && match != 0) {
          const int s = __ffs(match) - 1;
          idx = (((size_t)it * a.c_max + chain) * K + j) * S + s;
          if (lane == 0) {
            unsigned spins = 0;
            for (;;) {
              const unsigned stw = ld_acquire_u32(a.item_state + idx);
              if ((stw >> 2) != epoch) {  // nobody has started it: take it out of the workers' reach and solve it here
                if (atomicMax(a.item_state + idx, (epoch << 2) | 1u) < (epoch << 2)) {  // (states only grow: see item_state)
                  st_self += 1;
                  break;
                }
              } else if ((stw & 3u) == 2u) {
                how = 1;
                if (spins) st_wait += 1; else st_hit += 1;
                if (spins && a.timeline != nullptr) tl[5] += global_ns() - tl[7];
                break;
              } else {
                if (spins == 0 && a.timeline != nullptr) tl[7] = global_ns();
                ++spins;
                __nanosleep(40);
              }
            }
          }
          how = __shfl_sync(kFull, how, 0);
          __syncwarp();
        } else if (lane == 0 && j > 0) {
          st_miss += 1;
        }
        double cost, lamb_out;
        int iters;
        const size_t cbase = ((size_t)ep * L + lrel) * K;
        if (how == 1) {  // (table entries come from other SMs: L2 loads, the L1 of this SM may hold a stale line)
          lamb_out = __ldcg(a.item_lamb + idx);
          cost = __ldcg(a.item_cost + idx);
          iters = __ldcg(a.item_iters + idx);
          if (j == 0 || cost < best_cost) {  // first arg-min of the lap so far (utils/base.py:464): keep its trajectory
            br = j;
            best_cost = cost;
            const double* src = a.item_traj + idx * n_traj;
            double* xv = a.cand_x + (cbase + j) * 4 * (N + 1);
            double* uv = a.cand_u + (cbase + j) * 2 * N;
            for (int i = lane; i < 4 * (N + 1); i += 32) xv[i] = __ldcg(src + i);
            for (int i = lane; i < 2 * N; i += 32) uv[i] = __ldcg(src + 4 * (N + 1) + i);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) gs.c.xt[i] = lp.lap[(size_t)i * a.set.Tcap + jsel];
          gs.c.lamb = lamb;
          __syncwarp();
          gs.start(p);
          while (!gs.trip(p)) {
          }
          double xN[4];
          gs.terminal_state(xN);
          cost = relax_cost(p, xN, gs.c.xt, qfun, scale);
          lamb_out = gs.c.lamb;
          iters = gs.c.iters;
          if (j > 0) tl[6] += (1ull << 32) | (unsigned)iters;
          if (j == 0 || cost < best_cost) {
            br = j;
            best_cost = cost;
            gs.store(a.cand_x + (cbase + j) * 4 * (N + 1), a.cand_u + (cbase + j) * 2 * N);
          }
          __syncwarp();
        }
        if (lane == 0) {
          a.cand_cost[trow + j] = cost;
          a.cand_lamb[trow + j] = lamb_out;
          a.cand_iters[trow + j] = iters;
          st_iters += (unsigned long long)iters;
          st_solves += 1;
        }
        lamb = lamb_out;
        if (j == 0 && a.timeline != nullptr) tl[2] = global_ns();
      }
      (void)br;
      if (a.timeline != nullptr) tl[3] = global_ns();
      if (lane == 0) st_release_u64(walked, ((unsigned long long)epoch << 32) | 0xffffffffull);  // everything left is stale
      // ---- K3: the episode's last resolver to get here applies the winner rule (utils/base.py:462-478)
      __threadfence();
      __syncwarp();
      unsigned arrived = 0;
      if (lane == 0) arrived = atomicAdd(a.arrived + (size_t)it * a.e_stride + pos, 1u);
      arrived = __shfl_sync(kFull, arrived, 0);
      unsigned* done_flag = a.iter_done + (size_t)it * a.e_stride + pos;
      if ((int)arrived == L - 1) {
        __threadfence();
        WinnerArgs wa = a.win;
        wa.it = it;
        __syncwarp();
        winner_warp(wa, ep, lane, (size_t)L * K * sizeof(double) <= 16 * 1024 ? smem : nullptr);
        __threadfence();
        __syncwarp();
        if (it == kOuter - 1 && a.host_out != nullptr) {  // the step's result for this episode, straight into host memory
          double* ho = reinterpret_cast<double*>(a.host_out + 3 + (size_t)pos * spec_host_words(N));
          const int nx = 4 * (N + 1), nu = 2 * N;
          if (lane < 2) ho[lane] = a.win.u[(size_t)ep * 2 + lane];
          for (int i = lane; i < nx; i += 32) ho[2 + i] = a.win.x_pred[(size_t)ep * nx + i];
          for (int i = lane; i < nu; i += 32) ho[2 + nx + i] = a.win.u_pred[(size_t)ep * nu + i];
          if (lane == 0) {
            ho[2 + nx + nu] = a.win.winner_cost[ep];
            long long* hl = reinterpret_cast<long long*>(ho + 2 + nx + nu + 1);
            hl[0] = a.win.winner_idx[ep];
            hl[1] = ((long long)(unsigned)a.win.winner_rank[ep] << 32) | (unsigned)a.win.winner_lap[ep];
            hl[2] = (long long)a.win.shrink[ep];
          }
          __threadfence_system();
          __syncwarp();
        }
        if (lane == 0) st_release_u32(done_flag, epoch);
      } else {
        if (lane == 0)
          while (ld_acquire_u32(done_flag) != epoch) __nanosleep(40);
        __syncwarp();
      }
      __threadfence();
      if (a.timeline != nullptr && lane == 0) {
        tl[4] = global_ns();
        tl[7] = 0;
        for (int i = 0; i < 8; ++i) a.timeline[((size_t)it * a.c_max + chain) * 8 + i] = tl[i];
      }
    }
    // ---- end of the step for this chain: its share of the step's counters first, so that the last chain sees the totals
    if (lane == 0) {
      if (st_iters) atomicAdd(a.step_stats + 0, st_iters);
      if (st_solves) atomicAdd(a.step_stats + 1, st_solves);
      __threadfence();
    }
    st_iters = st_solves = 0;
    unsigned fin = 0;
    if (lane == 0) fin = atomicAdd(a.ctl + 2, 1u);
    fin = __shfl_sync(kFull, fin, 0);
    if ((int)fin == n_chains - 1 && a.host_out != nullptr && lane == 0) {
      __threadfence();
      volatile unsigned long long* ho = a.host_out;
      ho[1] = atomicAdd(a.step_stats + 0, 0ull);
      ho[2] = atomicAdd(a.step_stats + 1, 0ull);
      a.step_stats[0] = 0;  // (the no-copy form of the step has no separate reset launch)
      a.step_stats[1] = 0;
      __threadfence_system();
      ho[0] = (unsigned long long)epoch;  // the host is spinning on this word
      __threadfence_system();
    }
    if ((int)fin == n_chains - 1) {
      // last chain of the step: freeze the next launch's speculation set = the S most counted histogram entries
      __threadfence();
      unsigned c0 = a.hist_cnt[lane], c1 = a.hist_cnt[lane + 32];
      const unsigned long long k0 = a.hist_key[lane], k1 = a.hist_key[lane + 32];
      int r0 = 0, r1 = 0;
      unsigned cmax = 0;
      for (int src = 0; src < 32; ++src) {
        const unsigned o0 = __shfl_sync(kFull, c0, src), o1 = __shfl_sync(kFull, c1, src);
        r0 += (o0 > c0 || (o0 == c0 && src < lane)) + (o1 > c0);
        r1 += (o0 >= c1) + (o1 > c1 || (o1 == c1 && src < lane));
        cmax = o0 > cmax ? o0 : cmax;
        cmax = o1 > cmax ? o1 : cmax;
      }
      if (r0 < kSpecMaxSet) a.spec_set[r0] = k0;
      if (r1 < kSpecMaxSet) a.spec_set[r1] = k1;
      if (cmax > (1u << 20)) {  // keep the counts adaptive
        a.hist_cnt[lane] = c0 >> 1;
        a.hist_cnt[lane + 32] = c1 >> 1;
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release_u32(a.ctl + 3, epoch);  // workers leave
    }
  } else {
    // =============================================================== worker: speculative candidates, 8 per warp
    // Lane group g of all workers owns the table entries q = g, g + n_groups, g + 2 n_groups, ... of every outer
    // iteration, in the j-major order q = ((j - 1) * n_chains + chain) * S + s (per band of lambda values, see below), so
    // that the candidates a resolver reaches first are started first.  The groups of a warp run their candidates in lock step, so whatever one group does
    // between candidates is paid by all eight: picking the next entry is therefore fully pipelined.  While a candidate
    // runs, its group fetches the NEXT entry's published record in two asynchronous stages (stamped index word, then
    // terminal point / cost-to-go / resolver progress), each consumed a round after it was issued; when the candidate
    // ends the next one starts from registers, claimed by one atomicMax whose answer is read after the start-up
    // arithmetic.  No round waits for a global-memory round trip.
    constexpr int G = kWG;
    GroupSolver<double, kQR, kSpecRungs, G, S32> gs;
    gs.bind(smem, lane, N);
    const int rung = gs.rung;
    const unsigned gmask = (G >= 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u)) << gs.gbase_lane;
    auto bcast = [&](auto v) { return __shfl_sync(gmask, v, gs.gbase_lane); };
    enum { W_NEED, W_RUN, W_DONE };
    const unsigned long long n_groups = (unsigned long long)(gridDim.x - n_chains) * (32 / G);
    const unsigned long long gid = (unsigned long long)(ticket - n_chains) * (32 / G) + (lane / G);
    const int k_eff = a.k_eff < K ? a.k_eff : K;
    const unsigned long long per_it = (unsigned long long)n_chains * (k_eff > 1 ? k_eff - 1 : 0) * S;
    const int band = a.band > 0 && a.band < S ? a.band : S;
    const unsigned long long band_entries = (unsigned long long)n_chains * (k_eff > 1 ? k_eff - 1 : 0) * band;
    int state = per_it == 0 ? W_DONE : W_NEED;
    // running candidate
    int it = 0, j = 0, chain = 0;
    size_t idx = 0;
    double qfun = 0.0;
    // next entry of this group's list, being prefetched: 0 nothing issued, 1 index word in flight, 2 record in flight / here
    int pf_stage = 0, pf_it = 0, pf_j = 0, pf_chain = 0, pf_s = 0;
    unsigned long long pf_q = gid, pf_word = 0, pf_walked = 0;
    size_t pf_idx = 0;
    double pf_xt[4] = {0.0, 0.0, 0.0, 0.0}, pf_qfun = 0.0;
    bool list_done = per_it == 0;
    int bound_ep = -1;
    // claim of the running candidate whose answer has not been looked at yet; finished entry whose "done" flag is still to be set
    bool claim_pending = false, flag_pending = false;
    unsigned claim_prev = 0;
    size_t flag_idx = 0;
    unsigned over = 0;  // polled one round ahead: the load's latency hides under the round
    long long dbg_rounds = 0, dbg_pick = 0, dbg_trip = 0, dbg_pub = 0, dbg_t = 0;
    for (;;) {
      // ---- leave when the step is over (abandoning whatever is still in flight) or when nothing is left to try
      if (__shfl_sync(kFull, over, 0)) break;
      if (lane == 0) over = ld_relaxed_u32(a.ctl + 3) == epoch;
      if (__all_sync(kFull, state == W_DONE)) break;
      if (a.worker_dbg != nullptr) dbg_t = clock64();
      // ---- (z) left over from the previous round.  A finished entry's "done" flag is set one round after its data was
      // written (unless its resolver was already waiting for it): by then the stores have drained and the fence that must
      // precede the flag costs next to nothing, where a fence right behind the stores stalls all eight candidates of the warp.
      if (flag_pending) {
        __threadfence();
        __syncwarp(gmask);
        if (rung == 0) st_release_u32(a.item_state + flag_idx, (epoch << 2) | 2u);
        flag_pending = false;
      }
      if (claim_pending && state == W_RUN) {  // the claim's answer, a round after it was asked for
        claim_pending = false;
        if (bcast(claim_prev) >= (epoch << 2)) {  // somebody (the resolver, usually) already had this entry
          gs.running = false;
          state = W_NEED;
        }
      }
      // ---- (a) what was fetched during the previous round
      if (pf_stage == 1) {
        const unsigned long long w = bcast(pf_word);
        if ((w >> 32) == epoch) {
          pf_word = w;
          if ((unsigned)(w & 0xffffffffull) == 0xffffffffu || a.spec_set[pf_s] == kSpecUnused) {
            pf_stage = 0;  // the lap has no candidate j, or no lambda behind this slot: next entry
            pf_q += n_groups;
          } else {
            const double* rec = reinterpret_cast<const double*>(a.pub + (((size_t)pf_it * a.c_max + pf_chain) * K + pf_j) * kSpecPubWords);
#pragma unroll
            for (int i = 0; i < 4; ++i) pf_xt[i] = __ldcg(rec + 1 + i);
            pf_qfun = __ldcg(rec + 5);
            if (rung == 0) pf_walked = ld_relaxed_u64(a.walked + (size_t)pf_it * a.c_max + pf_chain);
            pf_stage = 2;
          }
        } else {
          pf_stage = 0;  // not published yet (the resolver is still selecting): ask again
        }
      }
      // ---- (b) start the prefetched candidate
      if (state == W_NEED && pf_stage == 2) {
        const unsigned long long wk = bcast(pf_walked);
        const bool stale = (wk >> 32) == epoch && (unsigned)(wk & 0xffffffffull) > (unsigned)pf_j;
        if (!stale) {
          it = pf_it;
          j = pf_j;
          chain = pf_chain;
          idx = pf_idx;
          qfun = pf_qfun;
          if (rung == 0) claim_prev = atomicMax(a.item_state + idx, (epoch << 2) | 1u);  // the claim; its answer is read a round later
          claim_pending = true;
          const int ep = a.ep_list != nullptr ? a.ep_list[chain / L] : chain / L;
          if (ep != bound_ep) {
            spec_bind_episode(p, a, ep, gs);
            bound_ep = ep;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) gs.c.xt[i] = pf_xt[i];
          gs.c.lamb = __longlong_as_double((long long)a.spec_set[pf_s]);
          gs.start(p);
          state = W_RUN;
          if (a.item_dbg != nullptr && rung == 0) {
            a.item_dbg[idx * 4 + 0] = global_ns();
            a.item_dbg[idx * 4 + 2] = 0;
          }
        }
        pf_stage = 0;
        pf_q += n_groups;
      }
      // ---- (c) issue the next fetch: the stamped index word of this group's next entry
      if (pf_stage == 0 && !list_done) {
        if (pf_q >= per_it) {  // this group's share of the outer iteration is done
          pf_q = gid;
          if (++pf_it == kOuter || pf_q >= per_it) list_done = true;
        }
        if (!list_done) {
          // two bands: the `band` most frequent lambda values of every (rank, chain) first, the rest of the set after them --
          // each band in the rank-major order above.  With more entries than lane groups (several waves) the likely hits of
          // ALL ranks are then under way in the first wave instead of only the first ranks' full sets.
          if (pf_q < band_entries) {
            pf_s = (int)(pf_q % band);
            pf_chain = (int)((pf_q / band) % n_chains);
            pf_j = 1 + (int)(pf_q / ((unsigned long long)band * n_chains));
          } else {
            const unsigned long long q2 = pf_q - band_entries;
            const int rest = S - band;
            pf_s = band + (int)(q2 % rest);
            pf_chain = (int)((q2 / rest) % n_chains);
            pf_j = 1 + (int)(q2 / ((unsigned long long)rest * n_chains));
          }
          pf_idx = (((size_t)pf_it * a.c_max + pf_chain) * K + pf_j) * S + pf_s;
          if (rung == 0) pf_word = ld_acquire_u64(a.pub + (((size_t)pf_it * a.c_max + pf_chain) * K + pf_j) * kSpecPubWords);
          pf_stage = 1;
        }
      }
      if (state == W_NEED && list_done && pf_stage == 0 && !flag_pending) state = W_DONE;
      const bool any_run = __any_sync(kFull, state == W_RUN);
      if (!any_run) {
        __nanosleep(60);
        continue;
      }
      // ---- one round of every live candidate of the warp
      unsigned long long wk = 0;
      if (state == W_RUN && rung == 0) wk = ld_relaxed_u64(a.walked + (size_t)it * a.c_max + chain);  // consumed after the round
      if (a.worker_dbg != nullptr) {
        const long long t = clock64();
        dbg_pick += t - dbg_t;
        dbg_t = t;
      }
      const bool fin = gs.trip(p);
      if (a.worker_dbg != nullptr) {
        const long long t = clock64();
        dbg_trip += t - dbg_t;
        dbg_t = t;
        dbg_rounds += 1;
      }
      if (a.item_dbg != nullptr && rung == 0 && state == W_RUN) a.item_dbg[idx * 4 + 2] += 1;
      if (fin) {
        bool mine = true;
        if (claim_pending) {  // a candidate that ended in its first round
          claim_pending = false;
          mine = bcast(claim_prev) < (epoch << 2);
        }
        if (mine) {
          if (a.item_dbg != nullptr && rung == 0) {
            a.item_dbg[idx * 4 + 1] = global_ns();
            a.item_dbg[idx * 4 + 3] = (unsigned long long)gs.c.iters;
          }
          double* tr = a.item_traj + idx * n_traj;
          gs.store(tr, tr + 4 * (N + 1));
          if (rung == 0) {
            double xN[4];
            gs.terminal_state(xN);
            a.item_cost[idx] = relax_cost(p, xN, gs.c.xt, qfun, it == 0 ? 1.0 : (it == 1 ? 10.0 : 100.0));
            a.item_lamb[idx] = gs.c.lamb;
            a.item_iters[idx] = gs.c.iters;
            st_spec += 1;
          }
          wk = bcast(wk);
          if ((wk >> 32) == epoch && (unsigned)(wk & 0xffffffffull) == (unsigned)j) {  // its resolver is waiting for it: now
            __threadfence();
            __syncwarp(gmask);
            if (rung == 0) st_release_u32(a.item_state + idx, (epoch << 2) | 2u);
          } else {
            flag_pending = true;
            flag_idx = idx;
          }
        }
        state = W_NEED;
      } else if (state == W_RUN) {
        wk = bcast(wk);
        if ((wk >> 32) == epoch && (unsigned)(wk & 0xffffffffull) > (unsigned)j) {  // the chain is past this candidate
          gs.running = false;
          if (rung == 0) st_drop += 1;
          state = W_NEED;
        }
      }
      __syncwarp();
      if (a.worker_dbg != nullptr) dbg_pub += clock64() - dbg_t;
    }
    if (a.worker_dbg != nullptr && lane == 0) {
      unsigned long long* d = a.worker_dbg + (size_t)blockIdx.x * 4;
      d[0] = (unsigned long long)dbg_rounds;
      d[1] = (unsigned long long)dbg_pick;
      d[2] = (unsigned long long)dbg_trip;
      d[3] = (unsigned long long)dbg_pub;
#ifdef I2LQR_TRIP_PROFILE
      unsigned long long* e = a.worker_dbg + 4096 + (size_t)blockIdx.x * 5;
      for (int i = 0; i < 5; ++i) e[i] = (unsigned long long)gs.prof[i];
#endif
    }
  }
  // ---- statistics, then the last block to leave restores the counters for the next launch
  const unsigned long long vals[8] = {st_iters, st_solves, st_hit, st_wait, st_self, st_miss, st_spec, st_drop};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unsigned long long v = vals[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if (lane == 0 && v != 0) atomicAdd(i < 2 ? a.step_stats + i : a.stats + (i - 2), v);
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    if (atomicAdd(a.ctl + 1, 1u) == gridDim.x - 1) {
      a.ctl[0] = 0;
      a.ctl[1] = 0;
      a.ctl[2] = 0;
      a.ctl[4] = a.ctl[5] = a.ctl[6] = 0;
      for (int i = 0; i < kOuter * a.e_stride; ++i) a.arrived[i] = 0;
      a.stats[6] += 1;
      __threadfence();
    }
  }
}

}  // namespace i2lqr
