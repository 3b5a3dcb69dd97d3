// i2lqr_api.cu -- C ABI (include/i2lqr.h) over the sm_100a kernels.  No CPU fallback: every compute entry
// point needs a CUDA device and fails with an error message otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/i2lqr.h"
#include "kernels.cuh"
#include "spec_step.cuh"
#include "components.cuh"
#include "kparams.h"

using namespace i2lqr;

static_assert(sizeof(DevObstacle) == sizeof(i2lqr_obstacle), "obstacle layouts must match");

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// Zeroing a few control words goes through a one-warp kernel, not cudaMemsetAsync: a memset may be queued on a
// copy engine and then waits behind another handle's bulk download, which stalls the solve launch behind it.
__global__ void zero_words_kernel(unsigned long long* p, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0ull;
}
// Same reason for the statistics readback: a 16-byte cudaMemcpy would queue behind another handle's bulk download,
// so a one-thread kernel writes the two counters straight into mapped pinned host memory.
__global__ void publish_words_kernel(const unsigned long long* src, volatile unsigned long long* dst_host, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst_host[i] = src[i];
  __threadfence_system();
}
inline cudaError_t zero_words(unsigned long long* p, int n, cudaStream_t st) {
  zero_words_kernel<<<1, 32, 0, st>>>(p, n);
  return cudaGetLastError();
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t want) {
    if (want <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) n = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

}  // namespace

struct i2lqr_handle {
  int device = 0;
  int n_sms = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t dev_done = nullptr;  // end of the last *_device call: the next one, on whatever stream, starts behind it
  i2lqr_params params;
  KParams kp;
  int E = 0, L = 0, K = 0, Tcap = 0;
  // statistics of the last compute call
  double last_ms = 0;
  long long last_launches = 0, last_iters = 0, last_solves = 0;
  // statistics words + work-fetch counters
  DevBuf<unsigned long long> ctl;
  int ctl_next = 0;
  static constexpr int kPipes = 8;      // streams of the chunked host pipeline (the handle's own + 7)
  cudaStream_t aux[kPipes - 1] = {};
  cudaEvent_t aux_done[kPipes - 1] = {};
  cudaEvent_t d2h_done = nullptr;        // all downloads of the last host-pointer batch have landed
  unsigned long long* h_stats = nullptr;  // mapped pinned host memory: statistics published by a kernel
  unsigned char* stage = nullptr;         // pinned staging block of the controller-step call
  size_t stage_bytes = 0;
  // uploads and downloads each get a stream of their own that never carries the other direction: the driver
  // binds a stream's copies to one copy engine, so a stream that uploads AND downloads queues both directions
  // on the same engine and the link never runs full duplex (measured: profiles/r01_host_pipeline_timeline.txt)
  cudaStream_t up = nullptr, down = nullptr;
  static constexpr int kMaxChunks = 16;
  cudaEvent_t ev_up[kMaxChunks] = {}, ev_solved[kMaxChunks] = {};
  DevBuf<double> gscratch[kPipes];
  // iteration-sliced passes: parked candidates (ping-pong), per pipeline stream
  DevBuf<double> park_state[kPipes][2];
  DevBuf<int> park_list[kPipes][2];
  DevBuf<unsigned long long> park_count[kPipes];  // one slab per pipeline stream (launches on different streams overlap)
  // flat batch staging
  DevBuf<double> b_x0, b_xt, b_lamb, b_uvar, b_xvar, b_lamb_out, b_uinit;
  DevBuf<DevObstacle> b_obs;
  DevBuf<int> b_iters;
  // safe sets
  SafeSet set{};
  std::vector<int> h_lap_len, h_lap_count;
  std::vector<DevObstacle> h_obs;
  DevBuf<DevObstacle> d_obs;
  bool obs_dirty = true;
  // obstacle lists (extension): up to kMaxExtraObstacles further ellipses per episode
  std::vector<DevObstacle> h_obs_extra;
  std::vector<int> h_n_extra;
  DevBuf<DevObstacle> d_obs_extra, d_obs_extra0;
  DevBuf<int> d_n_extra;
  int extras_in_use = 0;  // episodes with a list longer than one
  // controller-step workspace
  DevBuf<double> s_x, s_u, s_xpred, s_upred, s_cand_x, s_cand_u, s_cand_cost, s_cand_lamb, s_wcost;
  DevBuf<long long> s_sel_idx, s_widx, q_idx;
  bool split_open = false;  // between i2lqr_split_begin and i2lqr_split_end
  DevBuf<int> s_sel_cnt_own;
  DevBuf<int> s_sel_cnt, s_cand_iters, s_wlap, s_wrank, s_shrink, q_cnt;
  DevBuf<double> q_x0;
  // speculative one-launch controller step (spec_step.cuh)
  DevBuf<unsigned> sp_state, sp_ctl, sp_hist_cnt;
  DevBuf<double> sp_lamb, sp_cost, sp_traj;
  DevBuf<int> sp_iters;
  DevBuf<unsigned long long> sp_walked, sp_hist_key, sp_set, sp_stats, sp_timeline, sp_pub;
  int sp_cmax = 0, sp_K = 0, sp_S = 0, sp_E = 0, sp_N = 0;
  unsigned sp_epoch = 0;
  unsigned long long* sp_host = nullptr;      // mapped pinned: results of small steps, written by the kernel (spec_step.cuh host_out)
  unsigned long long* sp_host_dev = nullptr;  // its device alias
  size_t sp_host_words = 0;
  bool stats_lazy = false;  // last_ms of the no-copy step is read from the events on demand
  // component entry points (components.cuh): grow-only staging
  DevBuf<double> c_in[6], c_out[4], c_scratch;
  DevBuf<DevObstacle> c_obs;
  // batched episodes (K4)
  bool ep_ready = false;
  unsigned long long ep_noise_seed = 0;  // 0 = noise off (every reference script calls set_zero_noise)
  int ep_lap_index = 0;
  DevBuf<double> ep_upred, ep_goal, ep_lapx, ep_lapu;
  DevBuf<int> ep_step, ep_done, ep_horizon, ep_pb, ep_active, ep_status, ep_nrun, ep_laplen, ep_list, ep_nep;
  DevBuf<DevObstacle> d_obs0;
  int* h_nrun = nullptr;      // mapped pinned: [0] episodes still running, [1] episodes that solve the current step
  int* h_nrun_dev = nullptr;  // device alias of h_nrun
  static constexpr int kEpEvents = 8;
  cudaEvent_t ep_ev[kEpEvents] = {};
};

namespace {

int make_kparams(const i2lqr_params& p, KParams* k) {
  const char* err = fill_kparams(p, k);
  return err ? fail("%s", err) : 0;
}

constexpr size_t kSmemLimit = 227 * 1024;

constexpr int kCtlSlots = 16;   // fetch counters: one per launch that may be in flight concurrently
constexpr int kCtlWords = 2 + kCtlSlots;  // ctl[0] executed iterations, ctl[1] finished solves, then the counters
constexpr int kAuxStreams = i2lqr_handle::kPipes - 1;
constexpr long long kSliceMinBatch = 131072;  // smaller flat batches run as a single pass

// Launches the persistent solve kernel for `w` (counter / stats pointers are filled in here).
int launch_solve(i2lqr_handle* h, Work w, cudaStream_t st, int pipe = 0) {
  const KParams& kp = h->kp;
  // scratch placement: 0 shared memory (18N doubles per lane), 3 trajectory in shared memory + gains in tensor
  // memory (flat batches up to horizon 25: 12 / 8 / 4 resident warps per SM), 2 hybrid (8N in shared memory, 10N in
  // an L2-resident slab), 1 all in the slab (horizons whose trajectory does not fit shared memory)
  static const int force = getenv("I2LQR_SCRATCH") ? atoi(getenv("I2LQR_SCRATCH")) : -1;  // tuning switch
  const size_t smem_full = (size_t)kSolveThreads * 18 * kp.N * sizeof(double);
  const size_t smem_cur = (size_t)kSolveThreads * 8 * kp.N * sizeof(double);
  // (hybrid as long as ONE warp's trajectory fits: even at one warp per SM it beats the all-global placement, whose
  // stage loops wait on dependent L2 accesses -- N = 100: 1.79 vs 1.46 M solves/s)
  int scratch = smem_full * 4 <= kSmemLimit ? 0 : (smem_cur <= kSmemLimit ? 2 : 1);
  const bool tmem_ok = w.mode == 0 && !kp.has_qr && kp.N <= kTmemMaxHorizon && h->params.precision == 0;
  if (tmem_ok && w.n_items >= 8192) scratch = 3;
  // ... and up to horizon 6 with the candidate's cold state parked as well: 128 registers, four CTAs per SM
  static const bool no_lean = getenv("I2LQR_LEAN") && atoi(getenv("I2LQR_LEAN")) == 0;  // A-B switch
  static const bool lean_pass0_only = getenv("I2LQR_LEAN") && atoi(getenv("I2LQR_LEAN")) == 2;  // A-B switch
  // (all passes: mixing the 168-register and the 128-register kernel on an SM measured slower than either alone)
  if (scratch == 3 && kp.N <= kLeanMaxHorizon && !no_lean && (w.pass == 0 || !lean_pass0_only)) scratch = 6;
  // long horizons (flat batches, FP64): the working set is streamed through shared memory with bulk asynchronous copies
  const bool stream_ok = w.mode == 0 && h->params.precision == 0;
  static const int stream_from = getenv("I2LQR_STREAM_FROM") ? atoi(getenv("I2LQR_STREAM_FROM")) : 26;  // tuning switch: first horizon that streams
  if (stream_ok && kp.N >= stream_from && w.n_items >= 2048) scratch = 4;
  if (force >= 0) scratch = force;
  if ((scratch == 4 || scratch == 5) && !stream_ok) scratch = 2;
  if (scratch == 6 && !(tmem_ok && kp.N <= kLeanMaxHorizon)) scratch = 3;
  if (scratch == 3 && !tmem_ok) scratch = 0;
  if (scratch == 0 && smem_full > kSmemLimit) scratch = 2;
  if (scratch == 2 && smem_cur > kSmemLimit) scratch = 1;
  const int threads = (scratch == 3 || scratch == 6) ? kTmemThreads : kSolveThreads;
  CU(h->ctl.ensure(kCtlWords));
  const int slot = h->ctl_next;
  h->ctl_next = (slot + 1) % kCtlSlots;
  CU(zero_words(h->ctl.p + 2 + slot, 1, st));
  w.counter = h->ctl.p + 2 + slot;
  w.stats = h->ctl.p;
  long long blocks_needed = (w.n_items + threads - 1) / threads;
  if (blocks_needed < 1) blocks_needed = 1;
  const bool fp32 = h->params.precision == 1;
  if (fp32 && scratch >= 2) scratch = smem_full / 2 <= kSmemLimit ? 0 : 1;  // FP32 build: shared memory or global slab
  const size_t rsz = fp32 ? sizeof(float) : sizeof(double);
  using Kernel64 = void (*)(const KParams, const Work);
  using Kernel32 = void (*)(const KParamsT<float>, const Work);
  Kernel64 k64 = nullptr;
  Kernel32 k32 = nullptr;
  auto pick64 = [&](auto mode) -> Kernel64 {
    constexpr int M = decltype(mode)::value;
    constexpr int T = kSolveThreads;
    if (kp.has_qr) return scratch == 0 ? solve_kernel<double, true, 0, M, T> : (scratch == 1 ? solve_kernel<double, true, 1, M, T> : solve_kernel<double, true, 2, M, T>);
    return scratch == 0 ? solve_kernel<double, false, 0, M, T> : (scratch == 1 ? solve_kernel<double, false, 1, M, T> : solve_kernel<double, false, 2, M, T>);
  };
  auto pick32 = [&](auto mode) -> Kernel32 {
    constexpr int M = decltype(mode)::value;
    constexpr int T = kSolveThreads;
    if (kp.has_qr) return scratch == 0 ? solve_kernel<float, true, 0, M, T> : solve_kernel<float, true, 1, M, T>;
    return scratch == 0 ? solve_kernel<float, false, 0, M, T> : solve_kernel<float, false, 1, M, T>;
  };
  if (fp32) {
    k32 = w.mode == 0 ? pick32(std::integral_constant<int, 0>()) : pick32(std::integral_constant<int, 1>());
  } else {
    k64 = w.mode == 0 ? pick64(std::integral_constant<int, 0>()) : pick64(std::integral_constant<int, 1>());
    if (scratch == 3) k64 = solve_kernel<double, false, 3, 0, kTmemThreads>;
    if (scratch == 6) k64 = solve_kernel<double, false, 6, 0, kTmemThreads>;
    if (scratch == 4) k64 = kp.has_qr ? solve_kernel<double, true, 4, 0, kSolveThreads> : solve_kernel<double, false, 4, 0, kSolveThreads>;
    if (scratch == 5) k64 = kp.has_qr ? solve_kernel<double, true, 5, 0, kSolveThreads> : solve_kernel<double, false, 5, 0, kSolveThreads>;
  }
  const void* kern = fp32 ? (const void*)k32 : (const void*)k64;
  int per_sm = 0;
  size_t smem = scratch == 0 ? smem_full * rsz / sizeof(double)
                             : (scratch == 2 ? smem_cur : (scratch == 3 ? (size_t)kTmemThreads * 8 * kp.N * sizeof(double) : 0));
  if (scratch == 6) smem = (size_t)kTmemThreads * (8 * kp.N + 6) * sizeof(double);
  if ((scratch == 4 || scratch == 5)) smem = stream_smem_bytes(scratch == 4 ? kStreamRing : 4);
  // tuning switch for occupancy experiments: pad the dynamic shared memory so that at most this many blocks fit
  static const int cap_blocks = getenv("I2LQR_MAX_BLOCKS_PER_SM") ? atoi(getenv("I2LQR_MAX_BLOCKS_PER_SM")) : 0;
  if (smem > 0 && cap_blocks > 0) smem = std::max(smem, (size_t)(kSmemLimit / cap_blocks) - 1024);
  if (smem > 32 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (smem > 0) CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) return fail("solve kernel does not fit on an SM (horizon %d)", kp.N);
  // tensor memory: 512 columns per SM, 128 / 256 / 512 per CTA (horizon up to 6 / 12 / 25); registers (168 x 128
  // threads) allow three CTAs.  The occupancy calculator does not know about tensor memory, so the grid is sized by
  // hand: what the columns, the registers and the shared memory each allow.
  if (scratch == 3 || scratch == 6) {
    per_sm = std::min(scratch == 6 ? 4 : 3, 512 / tmem_columns(kp.N));
    per_sm = std::max(1, std::min<int>(per_sm, (int)((kSmemLimit + 1024) / (smem + 1024))));  // 228 KB per SM, 1 KB reserved per CTA
  }
  long long grid = std::min<long long>(blocks_needed, (long long)per_sm * h->n_sms);
  if ((scratch == 4 || scratch == 5)) {
    static const int stream_blocks = getenv("I2LQR_STREAM_BLOCKS") ? atoi(getenv("I2LQR_STREAM_BLOCKS")) : 0;  // tuning switch: warps per SM
    if (stream_blocks > 0) grid = std::min<long long>(grid, (long long)stream_blocks * h->n_sms);
    CU(h->gscratch[pipe].ensure((size_t)grid * stream_warp_doubles(kp.N)));
    w.gscratch = h->gscratch[pipe].p;
    w.gstride = 0;
  }
  if (scratch == 1 || scratch == 2) {
    const size_t per_lane = scratch == 1 ? 18 * kp.N : 10 * kp.N;
    CU(h->gscratch[pipe].ensure((size_t)grid * threads * per_lane));  // sized in doubles: enough for either precision
    w.gscratch = h->gscratch[pipe].p;
    w.gstride = grid * threads;
  }
  static const bool debug = getenv("I2LQR_DEBUG") != nullptr;
  if (debug) fprintf(stderr, "[i2lqr] solve launch: scratch %d grid %lld x %d smem %zu per_sm %d items %lld\n", scratch, grid, threads, smem, per_sm, w.n_items);
  if (fp32) {
    const KParamsT<float> kpf = to_float(kp);
    k32<<<(unsigned)grid, threads, smem, st>>>(kpf, w);
  } else {
    k64<<<(unsigned)grid, threads, smem, st>>>(kp, w);
  }
  CU(cudaGetLastError());
  h->last_launches += 1;
  return 0;
}

// Controller-step form: chain_kernel with kRungs speculative lambda rungs per chain.  The rung count follows the
// load: rungs trade idle lanes for fewer sequential trips, so they are used while every warp still has an SM
// sub-partition of its own and dropped once the chains alone fill the GPU.
template <typename R, int kRungs>
const void* pick_chain(bool has_qr, int scratch) {
  if (has_qr) return scratch == 0 ? (const void*)chain_kernel<R, true, 0, kRungs> : (scratch == 1 ? (const void*)chain_kernel<R, true, 1, kRungs> : (const void*)chain_kernel<R, true, 2, kRungs>);
  return scratch == 0 ? (const void*)chain_kernel<R, false, 0, kRungs> : (scratch == 1 ? (const void*)chain_kernel<R, false, 1, kRungs> : (const void*)chain_kernel<R, false, 2, kRungs>);
}

template <typename R, int kRungs, int kGroup>
const void* pick_chain_lat(bool has_qr) {
  return has_qr ? (const void*)chain_lat_kernel<R, true, kRungs, kGroup> : (const void*)chain_lat_kernel<R, false, kRungs, kGroup>;
}
template <typename R>
const void* pick_chain_lat(bool has_qr, int rungs, int group) {
  if (rungs == 4) return group == 32 ? pick_chain_lat<R, 4, 32>(has_qr) : (group == 8 ? pick_chain_lat<R, 4, 8>(has_qr) : pick_chain_lat<R, 4, 4>(has_qr));
  return rungs == 2 ? pick_chain_lat<R, 2, 2>(has_qr) : pick_chain_lat<R, 1, 1>(has_qr);
}

int launch_chain(i2lqr_handle* h, Work w, cudaStream_t st) {
  const KParams& kp = h->kp;
  const bool fp32 = h->params.precision == 1;
  const size_t rsz = fp32 ? sizeof(float) : sizeof(double);
  const size_t smem_full = (size_t)kSolveThreads * 18 * kp.N * rsz;
  const size_t smem_cur = (size_t)kSolveThreads * 8 * kp.N * rsz;
  static const int force = getenv("I2LQR_SCRATCH") ? atoi(getenv("I2LQR_SCRATCH")) : -1;  // tuning switch
  int scratch = smem_full * 4 <= kSmemLimit ? 0 : (smem_cur * 2 <= kSmemLimit ? 2 : 1);
  if (force >= 0 && force <= 2) scratch = force;
  if (scratch == 0 && smem_full > kSmemLimit) scratch = 2;
  if (scratch == 2 && smem_cur > kSmemLimit) scratch = 1;
  if (fp32 && scratch == 2) scratch = 1;
  static const int force_rungs = getenv("I2LQR_RUNGS") ? atoi(getenv("I2LQR_RUNGS")) : 0;  // tuning switch
  const long long slots = 4LL * h->n_sms;  // one warp per SM sub-partition
  // (four rungs up to TWO warps per sub-partition: at 8192 chains -- a 4096-episode lap at full load -- steps take 3.3-5.3 ms
  //  against 3.8-5.4 ms with two rungs in one warp per sub-partition; beyond that the rungs are dropped for lanes)
  int rungs = w.n_items * 4 <= 2 * slots * 32 ? 4 : (w.n_items * 2 <= slots * 32 ? 2 : 1);
  if (force_rungs == 1 || force_rungs == 2 || force_rungs == 4) rungs = force_rungs;
  const void* kern;
  if (fp32)
    kern = rungs == 4 ? pick_chain<float, 4>(kp.has_qr, scratch) : (rungs == 2 ? pick_chain<float, 2>(kp.has_qr, scratch) : pick_chain<float, 1>(kp.has_qr, scratch));
  else
    kern = rungs == 4 ? pick_chain<double, 4>(kp.has_qr, scratch) : (rungs == 2 ? pick_chain<double, 2>(kp.has_qr, scratch) : pick_chain<double, 1>(kp.has_qr, scratch));
  size_t smem = scratch == 0 ? smem_full : (scratch == 2 ? smem_cur : 0);
  // latency-oriented form (chain_core.cuh) whenever a warp's working set fits shared memory
  static const bool no_lat = getenv("I2LQR_CHAIN_LAT") && atoi(getenv("I2LQR_CHAIN_LAT")) == 0;  // tuning switch
  // lanes per chain: as many as keep the launch within one warp per SM sub-partition (the rungs first, then
  // extra lanes that only share the stage tasks)
  int group = rungs;
  if (rungs == 4) group = w.n_items <= slots ? 32 : (w.n_items * 8 <= slots * 32 ? 8 : 4);
  static const int force_group = getenv("I2LQR_GROUP") ? atoi(getenv("I2LQR_GROUP")) : 0;  // tuning switch
  if (rungs == 4 && (force_group == 4 || force_group == 8 || force_group == 32)) group = force_group;
  const size_t smem_lat = (size_t)kSolveThreads * chain_lane_elems(kp.N, group) * rsz;
  int lanes_per_chain = rungs;
  bool lat_form = false;
  if (!no_lat && smem_lat <= kSmemLimit) {
    lat_form = true;
    scratch = 0;
    smem = smem_lat;
    lanes_per_chain = group;
    kern = fp32 ? pick_chain_lat<float>(kp.has_qr, rungs, group) : pick_chain_lat<double>(kp.has_qr, rungs, group);
  }
  if (smem > 32 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (smem > 0) CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const long long lanes = w.n_items * lanes_per_chain;
  long long grid = std::max<long long>(1, (lanes + kSolveThreads - 1) / kSolveThreads);
  CU(h->ctl.ensure(kCtlWords));
  w.counter = nullptr;
  if (!lat_form && w.ep_n_extra != nullptr)
    return fail("obstacle lists need a horizon whose controller-step working set fits shared memory (num_horizon %d is too long)", kp.N);
  if (lat_form) {
    // latency form: never launch more warps than are resident at once -- with more chains than lane groups the groups
    // fetch further chains from a counter as they finish (a second wave of whole warps would double the time)
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSolveThreads, smem));
    const long long resident = (long long)std::max(1, per_sm) * h->n_sms;
    if (grid > resident) {
      grid = resident;
      const int slot = h->ctl_next;
      h->ctl_next = (slot + 1) % kCtlSlots;
      CU(zero_words(h->ctl.p + 2 + slot, 1, st));
      w.counter = h->ctl.p + 2 + slot;
    }
  }
  if (scratch != 0) {
    const size_t per_lane = scratch == 1 ? 18 * kp.N : 10 * kp.N;
    CU(h->gscratch[0].ensure((size_t)grid * kSolveThreads * per_lane));
    w.gscratch = h->gscratch[0].p;
    w.gstride = grid * kSolveThreads;
  }
  w.stats = h->ctl.p;
  static const bool debug = getenv("I2LQR_DEBUG") != nullptr;
  if (debug) fprintf(stderr, "[i2lqr] chain launch: rungs %d lanes/chain %d scratch %d grid %lld smem %zu chains %lld\n", rungs, lanes_per_chain, scratch, grid, smem, w.n_items);
  const KParamsT<float> kpf = fp32 ? to_float(kp) : KParamsT<float>{};
  void* args[2] = {fp32 ? (void*)&kpf : (void*)&kp, (void*)&w};
  CU(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(kSolveThreads), args, smem, st));
  h->last_launches += 1;
  return 0;
}

int begin_stats(i2lqr_handle* h, cudaStream_t st) {
  h->last_launches = 0;
  h->stats_lazy = false;
  CU(h->ctl.ensure(kCtlWords));
  CU(zero_words(h->ctl.p, 2, st));
  CU(cudaEventRecord(h->ev0, st));
  return 0;
}

// Waits for `st` and reads back the device-side counters and the ev0..ev1 kernel time.
int finish_stats(i2lqr_handle* h, cudaStream_t st) {
  if (!h->h_stats) CU(cudaHostAlloc((void**)&h->h_stats, 2 * sizeof(unsigned long long), cudaHostAllocMapped));
  publish_words_kernel<<<1, 32, 0, st>>>(h->ctl.p, h->h_stats, 2);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(st));
  const unsigned long long s[2] = {h->h_stats[0], h->h_stats[1]};
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_ms = ms;
  h->last_iters = (long long)s[0];
  h->last_solves = (long long)s[1];
  return 0;
}

int end_stats(i2lqr_handle* h, cudaStream_t st, bool sync) {
  CU(cudaEventRecord(h->ev1, st));
  return sync ? finish_stats(h, st) : 0;
}

int check_episode(i2lqr_handle* h, int e) {
  if (!h) return fail("null handle");
  if (e < 0 || e >= h->E) return fail("episode %d out of range [0,%d)", e, h->E);
  return 0;
}

int sync_obstacles(i2lqr_handle* h) {
  if (!h->obs_dirty) return 0;
  CU(h->d_obs.ensure(h->E));
  CU(cudaMemcpyAsync(h->d_obs.p, h->h_obs.data(), sizeof(DevObstacle) * h->E, cudaMemcpyHostToDevice, h->stream));
  if (h->extras_in_use > 0 || h->d_n_extra.p != nullptr) {
    CU(h->d_obs_extra.ensure((size_t)h->E * kMaxExtraObstacles));
    CU(h->d_n_extra.ensure(h->E));
    CU(cudaMemcpyAsync(h->d_obs_extra.p, h->h_obs_extra.data(), sizeof(DevObstacle) * h->E * kMaxExtraObstacles, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_n_extra.p, h->h_n_extra.data(), sizeof(int) * h->E, cudaMemcpyHostToDevice, h->stream));
  }
  h->obs_dirty = false;
  return 0;
}

int stage_reserve(i2lqr_handle* h, size_t bytes) {
  if (bytes <= h->stage_bytes) return 0;
  if (h->stage) cudaFreeHost(h->stage);
  h->stage = nullptr;
  h->stage_bytes = 0;
  CU(cudaMallocHost((void**)&h->stage, bytes));
  h->stage_bytes = bytes;
  return 0;
}

int knn_smem_config(i2lqr_handle* h, int* warps, size_t* smem) {
  // Tcap distances + Tcap ints (index array of the tie path), kept 8-byte aligned per warp
  const size_t per_warp = ((size_t)h->Tcap + ((size_t)h->Tcap + 1) / 2) * sizeof(double);
  int w = (int)std::min<size_t>(4, (200 * 1024) / per_warp);
  if (w < 1) return fail("lap_capacity %d too large for the K-nearest kernel", h->Tcap);
  *warps = w;
  *smem = per_warp * w;
  if (*smem > 48 * 1024) CU(cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
  return 0;
}

// Streams, events and device buffers of a new handle; on failure the caller destroys the half-built handle.
int create_resources(i2lqr_handle* h, const i2lqr_params* p, int n_episodes, int lap_capacity) {
  h->E = n_episodes;
  h->L = p->num_ss_iter;
  h->K = p->num_ss_points;
  h->Tcap = lap_capacity;
  CU(cudaSetDevice(h->device));
  // pipeline stream i carries chunk i of a host-pointer batch: earlier chunks get the higher priority so that their
  // last passes are scheduled ahead of later chunks' first passes and their download starts early
  int prio_least = 0, prio_greatest = 0;
  CU(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  static const bool no_prio = getenv("I2LQR_NO_PRIO") != nullptr;  // tuning switch
  auto pipe_priority = [&](int i) { return no_prio ? prio_least : std::min(prio_least, prio_greatest + i); };
  CU(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, pipe_priority(0)));
  CU(cudaEventCreate(&h->ev0));
  CU(cudaEventCreate(&h->ev1));
  CU(cudaEventCreateWithFlags(&h->d2h_done, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&h->dev_done, cudaEventDisableTiming));
  CU(cudaEventRecord(h->dev_done, h->stream));
  CU(cudaStreamCreateWithFlags(&h->up, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->down, cudaStreamNonBlocking));
  for (int i = 0; i < i2lqr_handle::kMaxChunks; ++i) {
    CU(cudaEventCreateWithFlags(&h->ev_up[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_solved[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < kAuxStreams; ++i) {
    CU(cudaStreamCreateWithPriority(&h->aux[i], cudaStreamNonBlocking, pipe_priority(1 + i)));
    CU(cudaEventCreateWithFlags(&h->aux_done[i], cudaEventDisableTiming));
  }
  CU(h->ctl.ensure(kCtlWords));
  // initialisation goes through the handle's own (non-blocking) stream: a legacy-stream cudaMemset is not
  // ordered against it and could land after the first add_lap / launch
  CU(cudaMemsetAsync(h->ctl.p, 0, kCtlWords * sizeof(unsigned long long), h->stream));
  if (n_episodes > 0) {
    const size_t ss_n = (size_t)h->E * h->L * 4 * h->Tcap;
    CU(cudaMalloc(&h->set.ss, ss_n * sizeof(double)));
    CU(cudaMemsetAsync(h->set.ss, 0, ss_n * sizeof(double), h->stream));
    CU(cudaMalloc(&h->set.lap_len, sizeof(int) * h->E * h->L));
    CU(cudaMalloc(&h->set.lap_count, sizeof(int) * h->E));
    CU(cudaMemsetAsync(h->set.lap_len, 0, sizeof(int) * h->E * h->L, h->stream));
    CU(cudaMemsetAsync(h->set.lap_count, 0, sizeof(int) * h->E, h->stream));
    h->set.L = h->L;
    h->set.Tcap = h->Tcap;
    h->h_lap_len.assign((size_t)h->E * h->L, 0);
    h->h_lap_count.assign(h->E, 0);
    DevObstacle none{};
    h->h_obs.assign(h->E, none);
    h->h_obs_extra.assign((size_t)h->E * kMaxExtraObstacles, none);
    h->h_n_extra.assign(h->E, 0);
  }
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

}  // namespace

extern "C" {

int i2lqr_abi_version(void) { return I2LQR_ABI_VERSION; }

const char* i2lqr_last_error(void) { return g_err.c_str(); }

void i2lqr_default_params(i2lqr_params* p) {
  memset(p, 0, sizeof(*p));
  p->num_horizon = 6;
  p->num_ss_points = 8;
  p->num_ss_iter = 1;
  p->max_ilqr_iter = 150;
  p->max_relax_iter = 55;
  p->max_sim_steps = 121;
  p->matrix_Qterminal[0] = 2.0;
  p->matrix_Qterminal[5] = 2.0;
  p->matrix_Qterminal[10] = 40.0;
  p->matrix_Qterminal[15] = 0.04;
  p->tuning_ctrl_q1 = p->tuning_ctrl_q2 = 1.0;
  p->tuning_obs_q1 = p->tuning_obs_q2 = 2.74;
  p->eps = 1e-2;
  p->lamb = 1.0;
  p->lamb_factor = 10.0;
  p->max_lamb = 1000.0;
  p->a_max = 2.0;
  p->delta_lim = 1.57;
  p->timestep = 1.0;
  p->goal_tol = 0.8;
}

int i2lqr_set_params(i2lqr_handle* h, const i2lqr_params* p) {
  if (!h || !p) return fail("null argument");
  if (h->E > 0 && (p->num_ss_iter != h->L)) return fail("num_ss_iter cannot change after create (safe-set ring size)");
  KParams kp;
  if (make_kparams(*p, &kp)) return 1;
  h->params = *p;
  h->kp = kp;
  h->K = p->num_ss_points;
  return 0;
}

int i2lqr_create(const i2lqr_params* p, int device, int n_episodes, int lap_capacity, i2lqr_handle** out) {
  if (!p || !out) return fail("null argument");
  if (p->precision != 0 && p->precision != 1) return fail("precision must be 0 (FP64) or 1 (FP32)");
  if (n_episodes < 0 || lap_capacity < 0) return fail("negative size");
  if (n_episodes > 0 && lap_capacity <= p->num_horizon) return fail("lap_capacity must exceed num_horizon");
  {
    KParams probe;  // every argument check happens before the first allocation
    if (make_kparams(*p, &probe)) return 1;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail("device %d out of range (have %d)", device, count);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail("device %d is sm_%d%d; this build targets sm_100a (B200)", device, prop.major, prop.minor);
  i2lqr_handle* h = new i2lqr_handle();
  h->device = device;
  h->n_sms = prop.multiProcessorCount;
  make_kparams(*p, &h->kp);
  h->params = *p;
  if (create_resources(h, p, n_episodes, lap_capacity)) {  // a CUDA failure half way: release what exists
    const std::string err = g_err;
    i2lqr_destroy(h);
    g_err = err;
    return 1;
  }
  *out = h;
  return 0;
}

int i2lqr_destroy(i2lqr_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto* b : {&h->b_x0, &h->b_xt, &h->b_lamb, &h->b_uinit, &h->b_uvar, &h->b_xvar, &h->b_lamb_out, &h->s_x, &h->s_u,
                  &h->s_xpred, &h->s_upred, &h->s_cand_x, &h->s_cand_u, &h->s_cand_cost, &h->s_cand_lamb, &h->s_wcost,
                  &h->q_x0})
    b->release();
  for (auto* b : {&h->s_sel_idx, &h->s_widx, &h->q_idx}) b->release();
  for (auto* b : {&h->b_iters, &h->s_sel_cnt, &h->s_sel_cnt_own, &h->s_cand_iters, &h->s_wlap, &h->s_wrank, &h->s_shrink, &h->q_cnt}) b->release();
  for (auto* b : {&h->ep_upred, &h->ep_goal, &h->ep_lapx, &h->ep_lapu}) b->release();
  for (auto* b : {&h->ep_step, &h->ep_done, &h->ep_horizon, &h->ep_pb, &h->ep_active, &h->ep_status, &h->ep_nrun, &h->ep_laplen, &h->ep_list, &h->ep_nep}) b->release();
  for (auto& e : h->ep_ev)
    if (e) cudaEventDestroy(e);
  h->d_obs0.release();
  h->d_obs_extra.release();
  h->d_obs_extra0.release();
  h->d_n_extra.release();
  for (auto* b : {&h->sp_state, &h->sp_ctl, &h->sp_hist_cnt}) b->release();
  for (auto* b : {&h->sp_lamb, &h->sp_cost, &h->sp_traj}) b->release();
  h->sp_iters.release();
  for (auto* b : {&h->sp_walked, &h->sp_hist_key, &h->sp_set, &h->sp_stats, &h->sp_timeline, &h->sp_pub}) b->release();
  for (auto& b : h->c_in) b.release();
  for (auto& b : h->c_out) b.release();
  h->c_scratch.release();
  h->c_obs.release();
  if (h->h_nrun) cudaFreeHost(h->h_nrun);
  if (h->h_stats) cudaFreeHost(h->h_stats);
  if (h->sp_host) cudaFreeHost(h->sp_host);
  if (h->stage) cudaFreeHost(h->stage);
  h->ctl.release();
  h->b_obs.release();
  h->d_obs.release();
  if (h->set.ss) cudaFree(h->set.ss);
  if (h->set.lap_len) cudaFree(h->set.lap_len);
  if (h->set.lap_count) cudaFree(h->set.lap_count);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->dev_done) cudaEventDestroy(h->dev_done);
  if (h->d2h_done) cudaEventDestroy(h->d2h_done);
  for (int i = 0; i < i2lqr_handle::kMaxChunks; ++i) {
    if (h->ev_up[i]) cudaEventDestroy(h->ev_up[i]);
    if (h->ev_solved[i]) cudaEventDestroy(h->ev_solved[i]);
  }
  if (h->up) cudaStreamDestroy(h->up);
  if (h->down) cudaStreamDestroy(h->down);
  for (int i = 0; i < i2lqr_handle::kPipes; ++i) {
    h->gscratch[i].release();
    h->park_count[i].release();
    for (int b = 0; b < 2; ++b) {
      h->park_state[i][b].release();
      h->park_list[i][b].release();
    }
  }
  for (int i = 0; i < kAuxStreams; ++i) {
    if (h->aux_done[i]) cudaEventDestroy(h->aux_done[i]);
    if (h->aux[i]) cudaStreamDestroy(h->aux[i]);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

// ------------------------------------------------------------------------------------------ flat batch
namespace {
int check_batch_args(i2lqr_handle* h, int64_t B, const double* x0, const double* x_term, const i2lqr_obstacle* obs,
                     int64_t n_obs) {
  if (!h) return fail("null handle");
  if (B < 0) return fail("negative batch");
  if (B > 0 && (!x0 || !x_term)) return fail("x0 and x_term are required");
  if (n_obs != 0 && n_obs != 1 && n_obs != B) return fail("n_obs must be 0, 1 or B");
  if (n_obs != 0 && !obs) return fail("obs is null but n_obs != 0");
  return 0;
}

// One launch over rows [lo, lo+n) of device-resident arrays.
int launch_flat(i2lqr_handle* h, int64_t lo, int64_t n, const double* x0, const double* x_term, const double* lamb_in,
                const i2lqr_obstacle* obs, int64_t n_obs, const double* uvar_init, double* uvar, double* xvar,
                double* lamb_out, int32_t* iters, cudaStream_t st, int pipe = 0) {
  const int N = h->kp.N;
  Work w{};
  w.mode = 0;
  w.n_items = n;
  w.x0 = x0 + lo * 4;
  w.x_term = x_term + lo * 4;
  w.lamb_in = lamb_in ? lamb_in + lo : nullptr;
  w.obs = reinterpret_cast<const DevObstacle*>(obs) + (n_obs > 1 ? lo : 0);
  w.n_obs = n_obs > 1 ? n : n_obs;
  w.u_init = uvar_init ? uvar_init + lo * 2 * N : nullptr;
  w.uvar = uvar ? uvar + lo * 2 * N : nullptr;
  w.xvar = xvar ? xvar + lo * 4 * (N + 1) : nullptr;
  w.lamb_out = lamb_out ? lamb_out + lo : nullptr;
  w.iters = iters ? iters + lo : nullptr;
  // Iteration slicing (see solve_kernel): big batches run as four passes -- candidates still unfinished after
  // 32 / 64 / 96 iterations are parked and resumed, packed into full warps, by the next pass.
  static const int cap0 = getenv("I2LQR_ITER_CAP") ? atoi(getenv("I2LQR_ITER_CAP")) : 32;  // tuning switch; 0 = off
  // (the streaming placement of long horizons runs as one pass: a resume pass costs up to 32 iterations x N stages of
  // latency whatever the number of parked candidates -- N = 200, 2^20 candidates: 251 ms in one pass, 280 ms in four)
  static const int stream_from_ = getenv("I2LQR_STREAM_FROM") ? atoi(getenv("I2LQR_STREAM_FROM")) : 26;
  const bool streams = h->params.precision == 0 && h->kp.N >= stream_from_ && !(getenv("I2LQR_SCRATCH") && atoi(getenv("I2LQR_SCRATCH")) < 4);
  if (cap0 <= 0 || n < kSliceMinBatch || streams) return launch_solve(h, w, st, pipe);
  const long long capacity = std::max<long long>(65536, n / 4);
  const size_t words = (size_t)(8 * N + 4);
  for (int b = 0; b < 2; ++b) {
    CU(h->park_state[pipe][b].ensure((size_t)capacity * words));
    CU(h->park_list[pipe][b].ensure((size_t)capacity));
  }
  CU(h->park_count[pipe].ensure(4));
  CU(zero_words(h->park_count[pipe].p, 4, st));
  const int n_pass = 4;
  for (int ps = 0; ps < n_pass; ++ps) {
    Work wp = w;
    wp.pass = ps;
    wp.iter_cap = ps + 1 < n_pass ? cap0 * (ps + 1) : 0;
    wp.park_capacity = capacity;
    wp.park_count = h->park_count[pipe].p + ps;
    wp.park_list = h->park_list[pipe][ps & 1].p;
    wp.park_state = h->park_state[pipe][ps & 1].p;
    if (ps > 0) {
      wp.resume_count = h->park_count[pipe].p + ps - 1;
      wp.resume_list = h->park_list[pipe][(ps - 1) & 1].p;
      wp.resume_state = h->park_state[pipe][(ps - 1) & 1].p;
      wp.n_items = capacity;  // upper bound for the grid; the kernel reads the real count on the device
    }
    if (launch_solve(h, wp, st, pipe)) return 1;
  }
  return 0;
}
}  // namespace

int i2lqr_solve_batch_device(i2lqr_handle* h, int64_t B, const double* x0, const double* x_term, const double* lamb_in,
                             const i2lqr_obstacle* obs, int64_t n_obs, const double* uvar_init, double* uvar,
                             double* xvar, double* lamb_out, int32_t* iters, void* cuda_stream) {
  if (check_batch_args(h, B, x0, x_term, obs, n_obs)) return 1;
  CU(cudaSetDevice(h->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  // A handle owns ONE set of scratch (parked candidates, the long-horizon slab, the statistics words): two calls on
  // the same handle must not overlap even when the caller passes different streams, so each call starts behind
  // the end of the previous one.  Callers that want overlapping batches use one handle per stream (bench.py).
  CU(cudaStreamWaitEvent(st, h->dev_done, 0));
  if (begin_stats(h, st)) return 1;
  if (B > 0 && launch_flat(h, 0, B, x0, x_term, lamb_in, obs, n_obs, uvar_init, uvar, xvar, lamb_out, iters, st)) return 1;
  if (end_stats(h, st, false)) return 1;
  CU(cudaEventRecord(h->dev_done, st));
  return 0;
}

// Host-pointer form.  Uploads (one stream that only ever uploads), solves (a stream per chunk) and downloads (one
// stream that only ever downloads) of up to four chunks overlap; with two or more caller threads, each on its
// own handle, one caller's upload + solve runs under the other's download and the call rate is bound by the
// D2H direction of PCIe (measured 51 GB/s of 57).  Nothing on this path uses a pageable copy: even the 16-byte
// statistics readback goes through mapped memory, because a pageable copy queues behind whatever bulk transfer
// another handle has in flight (that alone cost 30 % of the end-to-end rate, profiles/r01_host_pipeline.txt).
// The caller's buffers should be pinned; pageable buffers still work, the copies then serialise.
int i2lqr_solve_batch(i2lqr_handle* h, int64_t B, const double* x0, const double* x_term, const double* lamb_in,
                      const i2lqr_obstacle* obs, int64_t n_obs, const double* uvar_init, double* uvar, double* xvar,
                      double* lamb_out, int32_t* iters) {
  if (check_batch_args(h, B, x0, x_term, obs, n_obs)) return 1;
  CU(cudaSetDevice(h->device));
  const int N = h->kp.N;
  const size_t nb = (size_t)B;
  if (B > 0) {
    CU(h->b_x0.ensure(nb * 4));
    CU(h->b_xt.ensure(nb * 4));
    if (lamb_in) CU(h->b_lamb.ensure(nb));
    if (n_obs) CU(h->b_obs.ensure((size_t)n_obs));
    if (uvar_init) CU(h->b_uinit.ensure(nb * 2 * N));
    if (uvar) CU(h->b_uvar.ensure(nb * 2 * N));
    if (xvar) CU(h->b_xvar.ensure(nb * 4 * (N + 1)));
    if (lamb_out) CU(h->b_lamb_out.ensure(nb));
    if (iters) CU(h->b_iters.ensure(nb));
  }
  cudaStream_t main = h->stream;
  static const bool tl = getenv("I2LQR_TIMELINE") != nullptr;  // developer switch: per-chunk event timeline on stderr
  static cudaEvent_t tl_base = nullptr;
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t st) { if (tl) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
  std::vector<int64_t> sizes;      // chunk sizes
  {
  if (tl && !tl_base) {
    cudaEventCreate(&tl_base);
    cudaEventRecord(tl_base, main);
  }
  if (begin_stats(h, main)) return 1;
  if (B > 0) {
    if (n_obs == 1) CU(cudaMemcpyAsync(h->b_obs.p, obs, sizeof(DevObstacle), cudaMemcpyHostToDevice, main));
    CU(cudaEventRecord(h->ev0, main));  // aux streams start after the statistics reset (and the shared obstacle)
    cudaStream_t streams[1 + kAuxStreams];
    streams[0] = main;
    for (int i = 0; i < kAuxStreams; ++i) streams[1 + i] = h->aux[i];
    for (int i = 0; i < kAuxStreams; ++i) CU(cudaStreamWaitEvent(h->aux[i], h->ev0, 0));
    // Few, large chunks: a launch needs several candidates per lane to average out the 1..150 iterations a candidate
    // takes (2^17 candidates per launch run at 0.43 of the 2^20 rate), and iteration slicing needs >= 2^17.  Four
    // chunks let the first download start a quarter of the way in; finer cuts only cost GPU time.
    static const int max_chunks = getenv("I2LQR_CHUNKS") ? atoi(getenv("I2LQR_CHUNKS")) : 4;  // tuning switch
    const int64_t n_chunks = std::min<int64_t>(std::min(max_chunks, i2lqr_handle::kMaxChunks), std::max<int64_t>(1, B / (2 * kSliceMinBatch)));
    const int64_t chunk = (B + n_chunks - 1) / n_chunks;
    for (int64_t left = B; left > 0; left -= std::min(chunk, left)) sizes.push_back(std::min(chunk, left));
    const auto* d_obs = reinterpret_cast<const i2lqr_obstacle*>(h->b_obs.p);
    if ((int64_t)sizes.size() > i2lqr_handle::kMaxChunks) return fail("internal: too many chunks");
    CU(cudaStreamWaitEvent(h->up, h->ev0, 0));
    CU(cudaStreamWaitEvent(h->down, h->ev0, 0));
    // uploads of all chunks, in order, on the upload stream
    for (int64_t c = 0, lo = 0; c < (int64_t)sizes.size(); lo += sizes[c], ++c) {
      const size_t un = (size_t)sizes[c], ulo = (size_t)lo;
      cudaStream_t st = h->up;
      CU(cudaMemcpyAsync(h->b_x0.p + ulo * 4, x0 + ulo * 4, un * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(h->b_xt.p + ulo * 4, x_term + ulo * 4, un * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
      if (lamb_in) CU(cudaMemcpyAsync(h->b_lamb.p + ulo, lamb_in + ulo, un * sizeof(double), cudaMemcpyHostToDevice, st));
      if (n_obs > 1) CU(cudaMemcpyAsync(h->b_obs.p + ulo, obs + ulo, un * sizeof(DevObstacle), cudaMemcpyHostToDevice, st));
      if (uvar_init)
        CU(cudaMemcpyAsync(h->b_uinit.p + ulo * 2 * N, uvar_init + ulo * 2 * N, un * 2 * N * sizeof(double), cudaMemcpyHostToDevice, st));
      CU(cudaEventRecord(h->ev_up[c], st));
      mark(st);
    }
    // every chunk's solve on a compute stream of its own
    for (int64_t c = 0, lo = 0; c < (int64_t)sizes.size(); lo += sizes[c], ++c) {
      cudaStream_t st = streams[c % (1 + kAuxStreams)];
      CU(cudaStreamWaitEvent(st, h->ev_up[c], 0));
      if (launch_flat(h, lo, sizes[c], h->b_x0.p, h->b_xt.p, lamb_in ? h->b_lamb.p : nullptr, d_obs, n_obs,
                      uvar_init ? h->b_uinit.p : nullptr, uvar ? h->b_uvar.p : nullptr, xvar ? h->b_xvar.p : nullptr,
                      lamb_out ? h->b_lamb_out.p : nullptr, iters ? h->b_iters.p : nullptr, st, (int)(c % (1 + kAuxStreams))))
        return 1;
      CU(cudaEventRecord(h->ev_solved[c], st));
      mark(st);
    }
  }
  }
  if (B > 0) {
    // Downloads are only handed to the driver once they can run: a copy that still waits for its solve would sit
    // at the head of the copy engine's hardware queue and hold up every transfer queued behind it, including
    // other handles' (other caller threads') transfers.
    int64_t lo = 0;
    for (int64_t c = 0; c < (int64_t)sizes.size(); lo += sizes[c], ++c) {
      const size_t un = (size_t)sizes[c], ulo = (size_t)lo;
      cudaStream_t st = h->down;
      CU(cudaEventSynchronize(h->ev_solved[c]));
      if (uvar) CU(cudaMemcpyAsync(uvar + ulo * 2 * N, h->b_uvar.p + ulo * 2 * N, un * 2 * N * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (xvar)
        CU(cudaMemcpyAsync(xvar + ulo * 4 * (N + 1), h->b_xvar.p + ulo * 4 * (N + 1), un * 4 * (N + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (lamb_out) CU(cudaMemcpyAsync(lamb_out + ulo, h->b_lamb_out.p + ulo, un * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (iters) CU(cudaMemcpyAsync(iters + ulo, h->b_iters.p + ulo, un * sizeof(int), cudaMemcpyDeviceToHost, st));
      mark(st);
    }
    CU(cudaEventRecord(h->d2h_done, h->down));
    for (int i = 0; i < kAuxStreams; ++i) {
      CU(cudaEventRecord(h->aux_done[i], h->aux[i]));
      CU(cudaStreamWaitEvent(main, h->aux_done[i], 0));
    }
    CU(cudaStreamWaitEvent(main, h->d2h_done, 0));
  }
  CU(cudaEventRecord(h->ev1, main));  // here: whole pipeline (copies included), see i2lqr_last_stats
  const int rc = finish_stats(h, main);
  if (tl && rc == 0) {
    float t0 = 0;
    cudaEventElapsedTime(&t0, tl_base, h->ev0);
    const size_t nc = tev.size() / 3;
    for (size_t i = 0; i < nc; ++i) {
      float a, b, c2;
      cudaEventElapsedTime(&a, tl_base, tev[i]);
      cudaEventElapsedTime(&b, tl_base, tev[nc + i]);
      cudaEventElapsedTime(&c2, tl_base, tev[2 * nc + i]);
      fprintf(stderr, "[timeline %p] call@%.2f chunk %zu: uploaded %.2f  solved %.2f  downloaded %.2f ms\n", (void*)h, t0, i, a, b, c2);
    }
    for (auto e : tev) cudaEventDestroy(e);
  }
  return rc;
}

// -------------------------------------------------------------------------------------------- safe set
int i2lqr_add_lap(i2lqr_handle* h, int episode, const double* x, int T) {
  if (check_episode(h, episode)) return 1;
  if (!x) return fail("null lap");
  if (T <= h->kp.N) return fail("lap of %d points is not longer than the horizon %d (utils/base.py:348)", T, h->kp.N);
  if (T > h->Tcap) return fail("lap of %d points exceeds lap_capacity %d", T, h->Tcap);
  CU(cudaSetDevice(h->device));
  const int count = h->h_lap_count[episode];
  const int slot = count % h->L;
  std::vector<double> soa((size_t)4 * h->Tcap, 0.0);
  for (int j = 0; j < T; ++j)
    for (int i = 0; i < 4; ++i) soa[(size_t)i * h->Tcap + j] = x[(size_t)j * 4 + i];
  double* dst = h->set.ss + ((size_t)episode * h->L + slot) * 4 * (size_t)h->Tcap;
  CU(cudaMemcpyAsync(dst, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->h_lap_len[(size_t)episode * h->L + slot] = T;
  h->h_lap_count[episode] = count + 1;
  CU(cudaMemcpyAsync(h->set.lap_len + (size_t)episode * h->L + slot, &h->h_lap_len[(size_t)episode * h->L + slot],
                     sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->set.lap_count + episode, &h->h_lap_count[episode], sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));  // soa is a stack-lifetime staging buffer
  return 0;
}

namespace {
// One thread per (episode, point): the staged lap (T,4) is transposed into every episode's next ring slot.
__global__ void add_lap_all_kernel(SafeSet set, int E, const double* lap, int T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)E * set.Tcap) return;
  const int e = (int)(i / set.Tcap), j = (int)(i % set.Tcap);
  const int count = set.lap_count[e];
  const int slot = count % set.L;
  double* dst = set.ss + ((size_t)e * set.L + slot) * 4 * (size_t)set.Tcap;
#pragma unroll
  for (int c = 0; c < 4; ++c) dst[(size_t)c * set.Tcap + j] = j < T ? lap[(size_t)j * 4 + c] : 0.0;
}
__global__ void add_lap_all_commit_kernel(SafeSet set, int E, int T) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int count = set.lap_count[e];
  set.lap_len[(size_t)e * set.L + count % set.L] = T;
  set.lap_count[e] = count + 1;
}
}  // namespace

int i2lqr_add_lap_all(i2lqr_handle* h, const double* x, int T) {
  if (!h) return fail("null handle");
  if (h->E < 1) return fail("handle was created without episodes");
  if (!x) return fail("null lap");
  if (T <= h->kp.N) return fail("lap of %d points is not longer than the horizon %d (utils/base.py:348)", T, h->kp.N);
  if (T > h->Tcap) return fail("lap of %d points exceeds lap_capacity %d", T, h->Tcap);
  CU(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  CU(h->c_in[0].ensure((size_t)T * 4));
  if (stage_reserve(h, (size_t)T * 4 * sizeof(double))) return 1;
  memcpy(h->stage, x, (size_t)T * 4 * sizeof(double));
  CU(cudaMemcpyAsync(h->c_in[0].p, h->stage, (size_t)T * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
  const long long n = (long long)h->E * h->Tcap;
  add_lap_all_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->set, h->E, h->c_in[0].p, T);
  CU(cudaGetLastError());
  add_lap_all_commit_kernel<<<(h->E + 127) / 128, 128, 0, st>>>(h->set, h->E, T);
  CU(cudaGetLastError());
  for (int e = 0; e < h->E; ++e) {
    const int count = h->h_lap_count[e];
    h->h_lap_len[(size_t)e * h->L + count % h->L] = T;
    h->h_lap_count[e] = count + 1;
  }
  CU(cudaStreamSynchronize(st));  // the staging block is reused by the next call
  return 0;
}

int i2lqr_set_obstacles(i2lqr_handle* h, int first, int n, const i2lqr_obstacle* obs) {
  if (!h) return fail("null handle");
  if (first < 0 || n < 0 || first + n > h->E) return fail("episodes [%d, %d) out of range [0,%d)", first, first + n, h->E);
  for (int i = 0; i < n; ++i) {
    DevObstacle o{};
    if (obs) {
      memcpy(&o, obs + i, sizeof o);
      if (o.present && (o.moving_option < 0 || o.moving_option > 2)) return fail("episode %d: moving_option must be 0, 1 or 2", first + i);
      if (o.present && o.moving_option == 0 && o.spd != 0.0)
        return fail("episode %d: obstacle with spd != 0 needs moving_option 1 or 2 (control/ilqr_helper.py:34-43)", first + i);
      if (!o.present) o = DevObstacle{};
    }
    h->h_obs[first + i] = o;
    if (h->h_n_extra[first + i] > 0) {
      h->h_n_extra[first + i] = 0;
      h->extras_in_use -= 1;
    }
  }
  h->obs_dirty = true;
  return 0;
}

int i2lqr_num_laps(i2lqr_handle* h, int episode, int* n_laps) {
  if (check_episode(h, episode)) return 1;
  if (n_laps) *n_laps = h->h_lap_count[episode];
  return 0;
}

int i2lqr_set_obstacle(i2lqr_handle* h, int episode, const i2lqr_obstacle* obs) {
  if (check_episode(h, episode)) return 1;
  DevObstacle o{};
  if (obs) {
    memcpy(&o, obs, sizeof o);
    if (o.present && (o.moving_option < 0 || o.moving_option > 2)) return fail("moving_option must be 0, 1 or 2");
    if (o.present && o.moving_option == 0 && o.spd != 0.0)
      return fail("obstacle with spd != 0 needs moving_option 1 or 2 (control/ilqr_helper.py:34-43)");
  }
  if (h->h_n_extra[episode] > 0) {  // a single obstacle replaces a list
    h->h_n_extra[episode] = 0;
    h->extras_in_use -= 1;
    h->obs_dirty = true;
  }
  // (the drop-in controller passes its obstacle before every step: upload only what changed)
  if (memcmp(&h->h_obs[episode], &o, sizeof o) != 0) {
    h->h_obs[episode] = o;
    h->obs_dirty = true;
  }
  return 0;
}

int i2lqr_set_obstacle_list(i2lqr_handle* h, int episode, const i2lqr_obstacle* list, int n) {
  if (check_episode(h, episode)) return 1;
  if (n < 0 || n > 1 + kMaxExtraObstacles) return fail("an obstacle list holds 0..%d obstacles", 1 + kMaxExtraObstacles);
  if (n > 0 && !list) return fail("null obstacle list");
  if (n > 1 && h->params.precision != 0) return fail("obstacle lists are FP64 only");
  if (i2lqr_set_obstacle(h, episode, n > 0 ? list : nullptr)) return 1;  // validates and stores the first, clears an older list
  for (int m = 1; m < n; ++m) {
    DevObstacle o{};
    memcpy(&o, list + m, sizeof o);
    if (o.present && (o.moving_option < 0 || o.moving_option > 2)) return fail("moving_option must be 0, 1 or 2");
    if (o.present && o.moving_option == 0 && o.spd != 0.0)
      return fail("obstacle with spd != 0 needs moving_option 1 or 2 (control/ilqr_helper.py:34-43)");
    h->h_obs_extra[(size_t)episode * kMaxExtraObstacles + (m - 1)] = o;
  }
  if (n > 1) {
    h->h_n_extra[episode] = n - 1;
    h->extras_in_use += 1;
  }
  h->obs_dirty = true;
  return 0;
}

int i2lqr_select_knn(i2lqr_handle* h, int episode, int lap, const double* x0, int K, int64_t* idx_out, int32_t* n_out) {
  if (check_episode(h, episode)) return 1;
  if (!x0 || !idx_out) return fail("null argument");
  const int count = h->h_lap_count[episode];
  if (lap < 0 || lap >= count || lap < count - h->L) return fail("lap %d is not stored (have laps %d..%d)", lap, std::max(0, count - h->L), count - 1);
  if (K < 1) return fail("K must be >= 1");
  CU(cudaSetDevice(h->device));
  CU(h->q_x0.ensure(4));
  CU(h->q_idx.ensure((size_t)K));
  CU(h->q_cnt.ensure(1));
  CU(cudaMemcpyAsync(h->q_x0.p, x0, 4 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  KnnArgs a{};
  a.set = h->set;
  a.K = K;
  a.q_episode = episode;
  a.q_lap = lap;
  a.q_x0 = h->q_x0.p;
  a.q_idx = h->q_idx.p;
  a.q_cnt = h->q_cnt.p;
  int warps;
  size_t smem;
  if (knn_smem_config(h, &warps, &smem)) return 1;
  if (begin_stats(h, h->stream)) return 1;
  knn_kernel<<<1, 32 * warps, smem, h->stream>>>(a);
  CU(cudaGetLastError());
  h->last_launches += 1;
  int n = 0;
  CU(cudaMemcpyAsync(&n, h->q_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (end_stats(h, h->stream, true)) return 1;
  static_assert(sizeof(long long) == sizeof(int64_t), "");
  CU(cudaMemcpy(idx_out, h->q_idx.p, sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
  if (n_out) *n_out = n;
  return 0;
}

// ------------------------------------------------------------------------------------- controller step
namespace {
int ensure_step_workspace(i2lqr_handle* h) {
  const size_t E = h->E, L = h->L, K = h->K, N = h->kp.N;
  CU(h->s_x.ensure(E * 4));
  CU(h->s_u.ensure(E * 2));
  CU(h->s_xpred.ensure(E * 4 * (N + 1)));
  CU(h->s_upred.ensure(E * 2 * N));
  CU(h->s_cand_x.ensure(E * L * K * 4 * (N + 1)));
  CU(h->s_cand_u.ensure(E * L * K * 2 * N));
  CU(h->s_cand_cost.ensure(E * kOuter * L * K));
  CU(h->s_cand_lamb.ensure(E * kOuter * L * K));
  CU(h->s_cand_iters.ensure(E * kOuter * L * K));
  CU(h->s_sel_idx.ensure(E * kOuter * L * K));
  CU(h->s_sel_cnt.ensure(E * L));
  CU(h->s_wlap.ensure(E));
  CU(h->s_wrank.ensure(E));
  CU(h->s_widx.ensure(E));
  CU(h->s_wcost.ensure(E));
  CU(h->s_shrink.ensure(E));
  return 0;
}

// Which episodes a step's launches cover: the first n_episodes of the handle, or (batched closed loops) the
// device-side compacted list of the episodes that solve this step -- n_episodes is then only the host's upper
// bound of the list length, used to size the grids.
struct StepScope {
  int n_episodes;
  const int* active = nullptr;
  const int* ep_list = nullptr;
  const int* n_ep = nullptr;
};

// One outer iteration of calc_input (utils/base.py:384-478) as three launches; the split form of the step
// (i2lqr_split_*) calls them one by one with an exchange between K2 and K3.
int launch_knn_step(i2lqr_handle* h, int it, const StepScope& sc, cudaStream_t st) {
  const int n_episodes = sc.n_episodes;
  const int* d_active = sc.active;
  int warps;
  size_t smem;
  if (knn_smem_config(h, &warps, &smem)) return 1;
  KnnArgs ka{};
  ka.set = h->set;
  ka.K = h->K;
  ka.it = it;
  ka.n_episodes = n_episodes;
  ka.x = h->s_x.p;
  ka.x_pred = h->s_xpred.p;
  ka.N = h->kp.N;
  ka.active = d_active;
  ka.ep_list = sc.ep_list;
  ka.n_ep = sc.n_ep;
  ka.sel_idx = h->s_sel_idx.p;
  ka.sel_cnt = h->s_sel_cnt.p;
  const long long n_warps = (long long)n_episodes * h->L;
  knn_kernel<<<(unsigned)((n_warps + warps - 1) / warps), 32 * warps, smem, st>>>(ka);
  CU(cudaGetLastError());
  h->last_launches += 1;
  return 0;
}

// `sel_cnt` = candidates per (episode, lap) the chains walk: the K-nearest counts, or a masked copy of them
// when this handle only owns some of the laps.
int launch_chain_step(i2lqr_handle* h, int it, const StepScope& sc, const int* sel_cnt, cudaStream_t st) {
  const int* d_active = sc.active;
  Work w{};
  w.mode = 1;
  w.n_items = (long long)sc.n_episodes * h->L;
  w.ep_list = sc.ep_list;
  w.n_ep = sc.n_ep;
  w.it = it;
  w.scale = it == 0 ? 1.0 : (it == 1 ? 10.0 : 100.0);  // 10**it, utils/base.py:428-435
  w.set = h->set;
  w.x = h->s_x.p;
  w.ep_obs = h->d_obs.p;
  if (h->extras_in_use > 0) {
    w.ep_obs_extra = h->d_obs_extra.p;
    w.ep_n_extra = h->d_n_extra.p;
  }
  w.active = d_active;
  w.sel_idx = h->s_sel_idx.p;
  w.sel_cnt = sel_cnt;
  w.cand_x = h->s_cand_x.p;
  w.cand_u = h->s_cand_u.p;
  w.cand_cost = h->s_cand_cost.p;
  w.cand_lamb = h->s_cand_lamb.p;
  w.cand_iters = h->s_cand_iters.p;
  return launch_chain(h, w, st);
}

// fields of WinnerArgs that do not depend on the launch
void fill_winner_args(i2lqr_handle* h, WinnerArgs* out) {
  WinnerArgs wa{};
  wa.L = h->L;
  wa.K = h->K;
  wa.N = h->kp.N;
  wa.lap_count = h->set.lap_count;
  wa.lap_len = h->set.lap_len;
  wa.sel_cnt = h->s_sel_cnt.p;
  wa.sel_idx = h->s_sel_idx.p;
  wa.cand_cost = h->s_cand_cost.p;
  wa.cand_x = h->s_cand_x.p;
  wa.cand_u = h->s_cand_u.p;
  wa.u = h->s_u.p;
  wa.x_pred = h->s_xpred.p;
  wa.u_pred = h->s_upred.p;
  wa.winner_lap = h->s_wlap.p;
  wa.winner_rank = h->s_wrank.p;
  wa.winner_idx = h->s_widx.p;
  wa.winner_cost = h->s_wcost.p;
  wa.shrink = h->s_shrink.p;
  *out = wa;
}

int launch_winner_step(i2lqr_handle* h, int it, const StepScope& sc, cudaStream_t st) {
  const int n_episodes = sc.n_episodes;
  WinnerArgs wa;
  fill_winner_args(h, &wa);
  wa.ep_list = sc.ep_list;
  wa.n_ep = sc.n_ep;
  wa.n_episodes = n_episodes;
  wa.it = it;
  wa.active = sc.active;
  const int wthreads = 128;
  const size_t wsmem = (size_t)(wthreads / 32) * h->L * h->K * sizeof(double);
  wa.stage_costs = wsmem <= 48 * 1024;
  winner_kernel<<<(unsigned)(((long long)n_episodes * 32 + wthreads - 1) / wthreads), wthreads, wa.stage_costs ? wsmem : 0, st>>>(wa);
  CU(cudaGetLastError());
  h->last_launches += 1;
  return 0;
}

// unused trace slots read as "no candidate"
int clear_step_trace(i2lqr_handle* h, cudaStream_t st) {
  const size_t n = (size_t)h->E * kOuter * h->L * h->K;
  CU(cudaMemsetAsync(h->s_sel_idx.p, 0xff, sizeof(long long) * n, st));
  CU(cudaMemsetAsync(h->s_cand_cost.p, 0xff, sizeof(double) * n, st));
  CU(cudaMemsetAsync(h->s_cand_lamb.p, 0xff, sizeof(double) * n, st));
  CU(cudaMemsetAsync(h->s_cand_iters.p, 0xff, sizeof(int) * n, st));
  return 0;
}

// ---- speculative one-launch step (spec_step.cuh)
int spec_max_chains() {
  static const int n = getenv("I2LQR_SPEC_MAXCHAINS") ? std::max(1, atoi(getenv("I2LQR_SPEC_MAXCHAINS"))) : 256;  // tuning switch
  return n;
}
constexpr long long kSpecMaxEntries = 1ll << 19;  // table entries (344 B each at N = 6)

bool spec_enabled() {
  static const bool on = !(getenv("I2LQR_SPEC") && atoi(getenv("I2LQR_SPEC")) == 0);  // tuning / A-B switch
  return on;
}
int spec_set_size() {
  static const int s = getenv("I2LQR_SPEC_S") ? std::max(1, std::min(kSpecMaxSet, atoi(getenv("I2LQR_SPEC_S")))) : 32;  // tuning switch
  return s;
}
size_t spec_smem_bytes(const i2lqr_handle* h) {
  const int N = h->kp.N;
  const size_t solver = (size_t)kSolveThreads * std::max(chain_lane_elems(N, 32), chain_lane_elems(N, kSpecWorkerGroup)) * sizeof(double);
  const size_t knn = ((size_t)h->Tcap + ((size_t)h->Tcap + 1) / 2) * sizeof(double);
  return std::max(solver, knn);
}
// The one-launch form serves controller steps with few chains (one controller, a handful of episodes, the tail of a
// batched lap): FP64, chains that fit one warp's shared memory, a table of at most 2^16 speculative entries.
bool spec_eligible(const i2lqr_handle* h, int n_episodes) {
  if (!spec_enabled() || h->params.precision != 0) return false;
  const long long chains = (long long)n_episodes * h->L;
  if (chains < 1 || chains > spec_max_chains() || h->K < 2) return false;
  if (spec_smem_bytes(h) > 100 * 1024) return false;
  return (long long)kOuter * chains * h->K * kSpecMaxSet <= kSpecMaxEntries;
}

// lambda values reachable from lamb0 by the iteration's own /factor and *factor steps, breadth first (same
// operations, same doubles): the histogram's seed.  The first launches refine it from what the chains really see.
void spec_seed_patterns(const i2lqr_params& p, std::vector<unsigned long long>* keys) {
  std::vector<double> seen{p.lamb}, frontier{p.lamb};
  while (!frontier.empty() && (int)seen.size() < kSpecHist) {
    std::vector<double> next;
    for (double v : frontier)
      for (double w : {v * p.lamb_factor, v / p.lamb_factor}) {
        if (!std::isfinite(w) || w == 0.0 || w > p.max_lamb * p.lamb_factor * 1.5) continue;
        if (std::find(seen.begin(), seen.end(), w) != seen.end()) continue;
        seen.push_back(w);
        next.push_back(w);
      }
    frontier.swap(next);
  }
  keys->assign(kSpecHist, kSpecUnused);
  for (size_t i = 0; i < seen.size() && i < (size_t)kSpecHist; ++i) memcpy(&(*keys)[i], &seen[i], 8);
}

// The most episodes of this handle a step can have and still take the one-launch form (chain limit and table limit).
int spec_episode_cap(const i2lqr_handle* h) {
  const long long L = std::max(1, h->L), K = std::max(1, h->K);
  const long long by_chains = spec_max_chains() / L, by_table = kSpecMaxEntries / ((long long)kOuter * K * kSpecMaxSet) / L;
  return (int)std::max<long long>(1, std::min<long long>(h->E, std::min(by_chains, by_table)));
}

int spec_ensure(i2lqr_handle* h, int n_episodes, cudaStream_t st) {
  const int chains = n_episodes * h->L, S = kSpecMaxSet, K = h->K, N = h->kp.N;  // tables hold the largest set; a launch may use fewer
  const bool first = h->sp_hist_key.p == nullptr;
  if (first) {
    CU(h->sp_hist_key.ensure(kSpecHist));
    CU(h->sp_hist_cnt.ensure(kSpecHist));
    CU(h->sp_set.ensure(kSpecMaxSet));
    CU(h->sp_stats.ensure(16));  // [0..7] cumulative counters, [8..9] iterations / solves of a no-copy step (reset by the kernel)
    std::vector<unsigned long long> keys;
    spec_seed_patterns(h->params, &keys);
    std::vector<unsigned> cnt(kSpecHist, 0u);
    for (int i = 0; i < kSpecHist; ++i) cnt[i] = keys[i] == kSpecUnused ? 0u : (unsigned)(kSpecHist - i);  // breadth-first order = initial ranking
    CU(cudaMemcpyAsync(h->sp_hist_key.p, keys.data(), kSpecHist * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->sp_hist_cnt.p, cnt.data(), kSpecHist * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->sp_set.p, keys.data(), kSpecMaxSet * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(h->sp_stats.p, 0, 16 * 8, st));
    CU(cudaStreamSynchronize(st));  // the host vectors go out of scope
  }
  if (chains > h->sp_cmax || K != h->sp_K || S != h->sp_S || n_episodes > h->sp_E || N != h->sp_N) {
    // sized once for the largest step this handle can ever route here (its own episode count, capped by the form's limit):
    // the tail of a batched lap enters with a different episode count every lap, and growing the tables then means
    // cudaFree + cudaMalloc of ~70 MB in the middle of the lap (a device-wide synchronisation: measured as single steps of
    // up to 145 ms in a 4096-episode lap)
    const int e_cap = spec_episode_cap(h);
    const int cmax = std::max(std::max(chains, h->sp_cmax), e_cap * h->L), E = std::max(std::max(n_episodes, h->sp_E), e_cap);
    const size_t entries = (size_t)kOuter * cmax * K * S;
    CU(h->sp_state.ensure(entries));
    CU(h->sp_lamb.ensure(entries));
    CU(h->sp_cost.ensure(entries));
    CU(h->sp_iters.ensure(entries));
    CU(h->sp_traj.ensure(entries * (size_t)(6 * N + 4)));
    const size_t ctl_words = 8 + 2 * (size_t)kOuter * E;
    CU(h->sp_ctl.ensure(ctl_words));
    CU(h->sp_walked.ensure((size_t)kOuter * cmax));
    CU(h->sp_pub.ensure((size_t)kOuter * cmax * K * kSpecPubWords));
    CU(cudaMemsetAsync(h->sp_pub.p, 0, h->sp_pub.n * sizeof(unsigned long long), st));
    // a new layout invalidates every stamp: clear once (steady-state launches clear nothing)
    CU(cudaMemsetAsync(h->sp_state.p, 0, h->sp_state.n * sizeof(unsigned), st));
    CU(cudaMemsetAsync(h->sp_ctl.p, 0, h->sp_ctl.n * sizeof(unsigned), st));
    CU(cudaMemsetAsync(h->sp_walked.p, 0, h->sp_walked.n * sizeof(unsigned long long), st));
    h->sp_cmax = cmax;
    h->sp_E = E;
    h->sp_K = K;
    h->sp_S = S;
    h->sp_N = N;
  }
  return 0;
}

// The one-launch step in its two worker shapes: kSpecWorkerGroup (4) lanes per speculative candidate, eight candidates per
// worker warp -- or a whole warp per candidate, for steps whose table fits one wave of warps (one controller at the
// reference's default K = 8, L = 2: 448 entries; rounds of a candidate that shares no warp are a third shorter).
const void* spec_kernel(bool has_qr, int wg) {
  if (wg == 32) return has_qr ? (const void*)spec_step_kernel<true, 32> : (const void*)spec_step_kernel<false, 32>;
  return has_qr ? (const void*)spec_step_kernel<true> : (const void*)spec_step_kernel<false>;
}

// Everything the one-launch step needs the first time it runs -- its tables, the kernel's image (the runtime loads a
// function when it is first referenced) and its attributes -- done at set-up time, so that the first small step of a
// batched lap does not pay for it in the middle of the lap (measured: one step of 7 ... 41 ms).
int spec_prepare(i2lqr_handle* h, cudaStream_t st) {
  const int e_cap = spec_episode_cap(h);
  if (!spec_eligible(h, e_cap)) return 0;
  if (spec_ensure(h, e_cap, st)) return 1;
  const size_t smem = spec_smem_bytes(h);
  for (int wg : {kSpecWorkerGroup, 32}) {
    const void* kern = spec_kernel(h->kp.has_qr, wg);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSolveThreads, smem));
  }
  return 0;
}

int launch_spec_step(i2lqr_handle* h, const StepScope& sc, cudaStream_t st, const double* x_inline = nullptr, bool host_out = false) {
  if (spec_ensure(h, sc.n_episodes, st)) return 1;
  const KParams& kp = h->kp;
  SpecArgs a{};
  a.n_episodes = sc.n_episodes;
  a.e_stride = h->sp_E;
  a.ep_list = sc.ep_list;
  a.n_ep = sc.n_ep;
  a.set = h->set;
  a.x = h->s_x.p;
  if (x_inline != nullptr) {
    a.n_inline = sc.n_episodes;
    memcpy(a.x_inline, x_inline, sizeof(double) * 4 * sc.n_episodes);
  }
  if (host_out) {
    a.host_out = h->sp_host_dev;
    a.step_stats = h->sp_stats.p + 8;
  }
  a.ep_obs = h->d_obs.p;
  if (h->extras_in_use > 0) {
    a.ep_obs_extra = h->d_obs_extra.p;
    a.ep_n_extra = h->d_n_extra.p;
  }
  a.c_max = h->sp_cmax;
  a.sel_idx = h->s_sel_idx.p;
  a.sel_cnt = h->s_sel_cnt.p;
  a.cand_x = h->s_cand_x.p;
  a.cand_u = h->s_cand_u.p;
  a.cand_cost = h->s_cand_cost.p;
  a.cand_lamb = h->s_cand_lamb.p;
  a.cand_iters = h->s_cand_iters.p;
  fill_winner_args(h, &a.win);
  a.item_state = h->sp_state.p;
  a.item_lamb = h->sp_lamb.p;
  a.item_cost = h->sp_cost.p;
  a.item_iters = h->sp_iters.p;
  a.item_traj = h->sp_traj.p;
  if (h->sp_epoch >= (1u << 29)) {  // the stamps must keep growing: start over from a cleared table
    CU(cudaMemsetAsync(h->sp_state.p, 0, h->sp_state.n * sizeof(unsigned), st));
    CU(cudaMemsetAsync(h->sp_pub.p, 0, h->sp_pub.n * sizeof(unsigned long long), st));
    CU(cudaMemsetAsync(h->sp_walked.p, 0, h->sp_walked.n * sizeof(unsigned long long), st));
    h->sp_epoch = 0;
  }
  a.epoch = ++h->sp_epoch;
  a.ctl = h->sp_ctl.p;
  a.arrived = h->sp_ctl.p + 8;
  a.iter_done = a.arrived + (size_t)kOuter * h->sp_E;
  a.pub = h->sp_pub.p;
  {  // candidates per chain never exceed the longest stored lap (the work list of the workers is bounded by it)
    int longest = 0;
    for (int e = 0; e < h->E; ++e)
      for (int l = 0; l < h->L; ++l) longest = std::max(longest, h->h_lap_len[(size_t)e * h->L + l]);
    a.k_eff = std::max(1, std::min(h->K, longest));
  }
  a.walked = h->sp_walked.p;
  a.hist_key = h->sp_hist_key.p;
  a.hist_cnt = h->sp_hist_cnt.p;
  a.spec_set = h->sp_set.p;
  if (!host_out) a.step_stats = h->ctl.p;
  a.stats = h->sp_stats.p;
  static const bool want_timeline = getenv("I2LQR_SPEC_TIMELINE") != nullptr;  // developer switch
  if (want_timeline) {
    CU(h->sp_timeline.ensure((size_t)kOuter * h->sp_cmax * 8 + 4 * h->sp_state.n + 9 * 1024));
    CU(cudaMemsetAsync(h->sp_timeline.p, 0, h->sp_timeline.n * 8, st));
    a.timeline = h->sp_timeline.p;
    a.item_dbg = h->sp_timeline.p + (size_t)kOuter * h->sp_cmax * 8;
    a.worker_dbg = a.item_dbg + 4 * h->sp_state.n;  // (the grid never exceeds 1024 blocks)
  }
  // (lanes per speculative candidate: 4 and 8 measured the same step latency at the config-2 shape, 16 / 32 are 10-18 % slower:
  // fewer candidates in flight, more entries the resolver finds not started)
  const size_t smem = spec_smem_bytes(h);
  const long long chains = (long long)sc.n_episodes * h->L;
  // worker shape: a warp per speculative candidate while ONE outer iteration's entries fit one wave of warps, else eight per warp
  static const int force_wg = getenv("I2LQR_SPEC_WG") ? atoi(getenv("I2LQR_SPEC_WG")) : 0;  // A-B switch: 4 | 32
  int wg = kSpecWorkerGroup;
  {
    const void* k32 = spec_kernel(kp.has_qr, 32);
    CU(cudaFuncSetAttribute(k32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaFuncSetAttribute(k32, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm32 = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm32, k32, kSolveThreads, smem));
    const long long resident32 = (long long)per_sm32 * h->n_sms;
    if (chains * std::max(1, a.k_eff - 1) * spec_set_size() <= resident32 - chains) wg = 32;
  }
  if (force_wg == 4 || force_wg == 32) wg = force_wg == 4 ? kSpecWorkerGroup : 32;
  const void* kern = spec_kernel(kp.has_qr, wg);
  const int per_warp = 32 / wg;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSolveThreads, smem));
  if (per_sm < 1) return fail("speculative step kernel does not fit on an SM");
  const long long resident = (long long)per_sm * h->n_sms;
  if (resident <= chains) return fail("internal: no room for worker warps");
  // speculation set size: all of it while one wave of worker groups covers the table, halved down to 4 when the chains are many
  // (entries the workers cannot reach in time are solved in place by the resolvers: correct, just no faster than the chain itself)
  int S = spec_set_size();
  const long long group_slots = (resident - chains) * per_warp;
  static const int waves = getenv("I2LQR_SPEC_WAVES") ? std::max(1, atoi(getenv("I2LQR_SPEC_WAVES"))) : 8;  // tuning switch (2: 1.38 ms per step at 62 episodes, 8: 1.25)
  while (S > 4 && chains * std::max(1, a.k_eff - 1) * S > waves * group_slots) S /= 2;
  a.S = S;
  // (two bands only when the table needs more than two waves of lane groups -- 60 ... 128 episodes of a batched lap's tail:
  //  1.27 -> 1.18 ms per step at 62 episodes; at 1.4 waves, the K = 64, L = 8 controller, the plain rank-major order is
  //  5 % faster, and within one wave the order does not matter)
  static const int band_env = getenv("I2LQR_SPEC_BAND") ? atoi(getenv("I2LQR_SPEC_BAND")) : 8;  // tuning switch; 0 = one band
  a.band = (band_env > 0 && chains * std::max(1, a.k_eff - 1) * S > 2 * group_slots) ? band_env : 0;
  const long long workers = std::max<long long>(1, (chains * std::max(1, a.k_eff - 1) * S + per_warp - 1) / per_warp);
  const long long grid = std::min<long long>(resident, chains + workers);
  static const bool debug = getenv("I2LQR_DEBUG") != nullptr;
  if (debug) fprintf(stderr, "[i2lqr] spec step: chains %lld S %d lanes per candidate %d grid %lld smem %zu per_sm %d epoch %u\n", chains, a.S, wg, grid, smem, per_sm, a.epoch);
  void* args[2] = {(void*)&kp, (void*)&a};
  CU(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(kSolveThreads), args, smem, st));
  h->last_launches += 1;
  return 0;
}

// The three outer iterations of calc_input on device-resident state x (E,4); results stay on the device.
int run_step(i2lqr_handle* h, const StepScope& sc, cudaStream_t st, bool clear_trace = true) {
  if (sync_obstacles(h)) return 1;
  if (sc.active == nullptr && spec_eligible(h, sc.n_episodes)) return launch_spec_step(h, sc, st);  // one launch, no memsets
  if (clear_trace && clear_step_trace(h, st)) return 1;
  for (int it = 0; it < kOuter; ++it) {
    if (launch_knn_step(h, it, sc, st)) return 1;
    if (launch_chain_step(h, it, sc, h->s_sel_cnt.p, st)) return 1;
    if (launch_winner_step(h, it, sc, st)) return 1;
  }
  return 0;
}

// The caller's arrays are ordinary (pageable) memory: the state goes up and the results come down through one
// pinned staging block, so a controller-step call issues asynchronous copies only and blocks once.
size_t step_result_bytes(i2lqr_handle* h, int n_episodes) {
  const size_t E = n_episodes, N = h->kp.N;
  return 16 * 9 + E * (2 + 4 * (N + 1) + 2 * N + 1) * sizeof(double) + E * 3 * sizeof(int) + E * sizeof(long long);
}
// Results of the step that was just queued on `st`: device -> staging block (from byte `base`) -> caller.
int download_step_results(i2lqr_handle* h, int n_episodes, double* u, double* x_pred, double* u_pred, int32_t* winner_lap,
                          int32_t* winner_rank, int64_t* winner_idx, double* winner_cost, int32_t* shrink, size_t base,
                          cudaStream_t st) {
  const size_t E = n_episodes, N = h->kp.N;
  const size_t sz[8] = {E * 2 * sizeof(double), E * 4 * (N + 1) * sizeof(double), E * 2 * N * sizeof(double), E * sizeof(int),
                        E * sizeof(int),        E * sizeof(long long),            E * sizeof(double),         E * sizeof(int)};
  const void* src[8] = {h->s_u.p, h->s_xpred.p, h->s_upred.p, h->s_wlap.p, h->s_wrank.p, h->s_widx.p, h->s_wcost.p, h->s_shrink.p};
  void* dst[8] = {u, x_pred, u_pred, winner_lap, winner_rank, winner_idx, winner_cost, shrink};
  size_t off[8], total = base;
  for (int i = 0; i < 8; ++i) {
    off[i] = total;
    total += (sz[i] + 15) / 16 * 16;
  }
  if (total > h->stage_bytes) return fail("staging block too small (%zu > %zu)", total, h->stage_bytes);
  for (int i = 0; i < 8; ++i)
    if (dst[i]) CU(cudaMemcpyAsync(h->stage + off[i], src[i], sz[i], cudaMemcpyDeviceToHost, st));
  if (finish_stats(h, st)) return 1;
  for (int i = 0; i < 8; ++i)
    if (dst[i]) memcpy(dst[i], h->stage + off[i], sz[i]);
  return 0;
}
}  // namespace

namespace {
// The no-copy form of a small controller step: the state travels in the launch parameters, the kernel writes the results
// into mapped host memory and raises a flag the host spins on.  One launch, no copy, no stream synchronisation; the
// kernel's tail (workers leaving, counters restored) overlaps the caller's next host work, and the next launch queues
// behind it on the stream.
int solve_step_nocopy(i2lqr_handle* h, int n_episodes, const double* x, double* u, double* x_pred, double* u_pred,
                      int32_t* winner_lap, int32_t* winner_rank, int64_t* winner_idx, double* winner_cost, int32_t* shrink) {
  cudaStream_t st = h->stream;
  const int N = h->kp.N;
  const size_t words = 3 + (size_t)kSpecInlineEpisodes * spec_host_words(N);
  if (h->sp_host_words < words) {
    if (h->sp_host) CU(cudaFreeHost(h->sp_host));
    h->sp_host = nullptr;
    CU(cudaHostAlloc((void**)&h->sp_host, words * sizeof(unsigned long long), cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer((void**)&h->sp_host_dev, h->sp_host, 0));
    memset(h->sp_host, 0, words * sizeof(unsigned long long));
    h->sp_host_words = words;
  }
  if (sync_obstacles(h)) return 1;
  h->last_launches = 0;
  CU(cudaEventRecord(h->ev0, st));
  if (launch_spec_step(h, StepScope{n_episodes}, st, x, true)) return 1;
  CU(cudaEventRecord(h->ev1, st));
  volatile unsigned long long* ho = h->sp_host;
  const unsigned long long want = h->sp_epoch;
  for (unsigned long long spins = 0; ho[0] != want; ++spins) {
    if ((spins & 0xfffff) == 0xfffff) {  // ~ every few ms: make sure the kernel is still alive
      const cudaError_t q = cudaStreamQuery(st);
      if (q != cudaErrorNotReady && ho[0] != want) {
        if (q != cudaSuccess) return fail("controller step failed: %s", cudaGetErrorString(q));
        return fail("controller step finished without publishing its result");
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  h->last_iters = (long long)ho[1];
  h->last_solves = (long long)ho[2];
  h->stats_lazy = true;
  const int nx = 4 * (N + 1), nu = 2 * N;
  for (int e = 0; e < n_episodes; ++e) {
    const double* r = reinterpret_cast<const double*>(h->sp_host + 3 + (size_t)e * spec_host_words(N));
    const long long* rl = reinterpret_cast<const long long*>(r + 2 + nx + nu + 1);
    if (u) memcpy(u + (size_t)e * 2, r, 2 * sizeof(double));
    if (x_pred) memcpy(x_pred + (size_t)e * nx, r + 2, nx * sizeof(double));
    if (u_pred) memcpy(u_pred + (size_t)e * nu, r + 2 + nx, nu * sizeof(double));
    if (winner_cost) winner_cost[e] = r[2 + nx + nu];
    if (winner_idx) winner_idx[e] = rl[0];
    if (winner_lap) winner_lap[e] = (int32_t)(rl[1] & 0xffffffffll);
    if (winner_rank) winner_rank[e] = (int32_t)(rl[1] >> 32);
    if (shrink) shrink[e] = (int32_t)rl[2];
  }
  return 0;
}
}  // namespace

int i2lqr_solve_step(i2lqr_handle* h, int n_episodes, const double* x, double* u, double* x_pred, double* u_pred,
                     int32_t* winner_lap, int32_t* winner_rank, int64_t* winner_idx, double* winner_cost,
                     int32_t* shrink) {
  if (!h) return fail("null handle");
  if (n_episodes < 1 || n_episodes > h->E) return fail("n_episodes %d out of range [1,%d]", n_episodes, h->E);
  if (!x) return fail("null state");
  for (int e = 0; e < n_episodes; ++e)
    if (h->h_lap_count[e] == 0) return fail("episode %d has no stored lap (call add_trajectory first)", e);
  CU(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  if (ensure_step_workspace(h)) return 1;
  static const bool no_nocopy = getenv("I2LQR_STEP_COPY") != nullptr;  // A-B switch: force the staged-copy form
  static const bool want_timeline = getenv("I2LQR_SPEC_TIMELINE") != nullptr;
  if (!no_nocopy && !want_timeline && n_episodes <= kSpecInlineEpisodes && spec_eligible(h, n_episodes))
    return solve_step_nocopy(h, n_episodes, x, u, x_pred, u_pred, winner_lap, winner_rank, winner_idx, winner_cost, shrink);
  h->stats_lazy = false;
  const size_t n_in = (size_t)n_episodes * 4 * sizeof(double);
  if (stage_reserve(h, n_in + step_result_bytes(h, n_episodes))) return 1;
  memcpy(h->stage, x, n_in);
  CU(cudaMemcpyAsync(h->s_x.p, h->stage, n_in, cudaMemcpyHostToDevice, st));
  if (begin_stats(h, st)) return 1;
  if (run_step(h, StepScope{n_episodes}, st)) return 1;
  CU(cudaEventRecord(h->ev1, st));
  return download_step_results(h, n_episodes, u, x_pred, u_pred, winner_lap, winner_rank, winner_idx, winner_cost, shrink,
                               (n_in + 15) / 16 * 16, st);
}

int i2lqr_get_step_trace(i2lqr_handle* h, int n_episodes, int64_t* sel_idx, double* cand_cost, double* cand_lamb,
                         int32_t* cand_iters) {
  if (!h) return fail("null handle");
  if (n_episodes < 1 || n_episodes > h->E) return fail("n_episodes out of range");
  if (!h->s_sel_idx.p) return fail("no controller step has run yet");
  CU(cudaSetDevice(h->device));
  const size_t n = (size_t)n_episodes * kOuter * h->L * h->K;
  if (sel_idx) CU(cudaMemcpy(sel_idx, h->s_sel_idx.p, n * sizeof(long long), cudaMemcpyDeviceToHost));
  if (cand_cost) CU(cudaMemcpy(cand_cost, h->s_cand_cost.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (cand_lamb) CU(cudaMemcpy(cand_lamb, h->s_cand_lamb.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (cand_iters) CU(cudaMemcpy(cand_iters, h->s_cand_iters.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

// --------------------------------------------------------------------- controller step split over handles
namespace {
int split_check(i2lqr_handle* h, int n_episodes, int it, int part, int n_parts) {
  if (!h) return fail("null handle");
  if (n_episodes < 1 || n_episodes > h->E) return fail("n_episodes %d out of range [1,%d]", n_episodes, h->E);
  if (it < 0 || it >= kOuter) return fail("outer iteration %d out of range [0,%d)", it, kOuter);
  if (n_parts < 1 || part < 0 || part >= n_parts) return fail("part %d of %d is not a valid split", part, n_parts);
  if (!h->split_open) return fail("no split step is open (call i2lqr_split_begin first)");
  return 0;
}
SplitArgs split_args(i2lqr_handle* h, int n_episodes, int it, int part, int n_parts, double* rows) {
  SplitArgs a{};
  a.n_episodes = n_episodes;
  a.it = it;
  a.L = h->L;
  a.K = h->K;
  a.N = h->kp.N;
  a.part = part;
  a.n_parts = n_parts;
  a.lap_count = h->set.lap_count;
  a.sel_cnt = h->s_sel_cnt.p;
  a.sel_cnt_own = h->s_sel_cnt_own.p;
  a.cand_cost = h->s_cand_cost.p;
  a.cand_lamb = h->s_cand_lamb.p;
  a.cand_iters = h->s_cand_iters.p;
  a.cand_x = h->s_cand_x.p;
  a.cand_u = h->s_cand_u.p;
  a.rows = rows;
  return a;
}
}  // namespace

int i2lqr_split_row_doubles(i2lqr_handle* h, int64_t* row_doubles) {
  if (!h || !row_doubles) return fail("null argument");
  *row_doubles = split_row_doubles(h->K, h->kp.N);
  return 0;
}

int i2lqr_split_begin(i2lqr_handle* h, int n_episodes, const double* x, void* cuda_stream) {
  if (!h) return fail("null handle");
  if (n_episodes < 1 || n_episodes > h->E) return fail("n_episodes %d out of range [1,%d]", n_episodes, h->E);
  if (!x) return fail("null state");
  for (int e = 0; e < n_episodes; ++e)
    if (h->h_lap_count[e] == 0) return fail("episode %d has no stored lap (call add_trajectory first)", e);
  CU(cudaSetDevice(h->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  if (ensure_step_workspace(h)) return 1;
  CU(h->s_sel_cnt_own.ensure((size_t)h->E * h->L));
  if (sync_obstacles(h)) return 1;
  if (st != h->stream) CU(cudaStreamSynchronize(h->stream));  // the obstacle upload rides the handle's stream
  const size_t n_in = (size_t)n_episodes * 4 * sizeof(double);
  if (stage_reserve(h, n_in + step_result_bytes(h, n_episodes))) return 1;
  memcpy(h->stage, x, n_in);
  CU(cudaMemcpyAsync(h->s_x.p, h->stage, n_in, cudaMemcpyHostToDevice, st));
  if (begin_stats(h, st)) return 1;
  if (clear_step_trace(h, st)) return 1;
  h->split_open = true;
  return 0;
}

int i2lqr_split_solve(i2lqr_handle* h, int n_episodes, int it, int part, int n_parts, double* rows_device,
                      void* cuda_stream) {
  if (split_check(h, n_episodes, it, part, n_parts)) return 1;
  if (!rows_device) return fail("null exchange buffer");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  if (launch_knn_step(h, it, StepScope{n_episodes}, st)) return 1;
  const SplitArgs a = split_args(h, n_episodes, it, part, n_parts, rows_device);
  const int n_laps = n_episodes * h->L;
  split_mask_kernel<<<(n_laps + 127) / 128, 128, 0, st>>>(a);
  CU(cudaGetLastError());
  if (launch_chain_step(h, it, StepScope{n_episodes}, h->s_sel_cnt_own.p, st)) return 1;
  split_pack_kernel<<<(n_laps * 32 + 127) / 128, 128, 0, st>>>(a);
  CU(cudaGetLastError());
  h->last_launches += 2;
  return 0;
}

int i2lqr_split_merge(i2lqr_handle* h, int n_episodes, int it, int part, int n_parts, const double* all_rows_device,
                      void* cuda_stream) {
  if (split_check(h, n_episodes, it, part, n_parts)) return 1;
  if (!all_rows_device) return fail("null exchange buffer");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  const SplitArgs a = split_args(h, n_episodes, it, part, n_parts, const_cast<double*>(all_rows_device));
  const int n_laps = n_episodes * h->L;
  split_unpack_kernel<<<(n_laps * 32 + 127) / 128, 128, 0, st>>>(a);
  CU(cudaGetLastError());
  h->last_launches += 1;
  return launch_winner_step(h, it, StepScope{n_episodes}, st);
}

int i2lqr_split_end(i2lqr_handle* h, int n_episodes, double* u, double* x_pred, double* u_pred, int32_t* winner_lap,
                    int32_t* winner_rank, int64_t* winner_idx, double* winner_cost, int32_t* shrink, void* cuda_stream) {
  if (!h) return fail("null handle");
  if (n_episodes < 1 || n_episodes > h->E) return fail("n_episodes %d out of range [1,%d]", n_episodes, h->E);
  if (!h->split_open) return fail("no split step is open (call i2lqr_split_begin first)");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  h->split_open = false;
  CU(cudaEventRecord(h->ev1, st));
  if (download_step_results(h, n_episodes, u, x_pred, u_pred, winner_lap, winner_rank, winner_idx, winner_cost, shrink, 0, st))
    return 1;
  return 0;
}

// ------------------------------------------------------------------------------ component entry points
namespace {
struct CompIO {  // host arrays of one component call: inputs go up, outputs come down, all on the handle's stream
  i2lqr_handle* h;
  cudaStream_t st;
  int n_in = 0, n_out = 0;
  double* host_out[4] = {};
  size_t out_count[4] = {};
  int in(const double* src, size_t count, const double** dev) {
    *dev = nullptr;
    if (!src) return 0;
    DevBuf<double>& b = h->c_in[n_in++];
    CU(b.ensure(count));
    CU(cudaMemcpyAsync(b.p, src, count * sizeof(double), cudaMemcpyHostToDevice, st));
    *dev = b.p;
    return 0;
  }
  int out(double* dst, size_t count, double** dev) {
    *dev = nullptr;
    if (!dst) return 0;
    DevBuf<double>& b = h->c_out[n_out];
    CU(b.ensure(count));
    host_out[n_out] = dst;
    out_count[n_out] = count;
    ++n_out;
    *dev = b.p;
    return 0;
  }
  int obstacles(const i2lqr_obstacle* obs, int64_t n_obs, const DevObstacle** dev) {
    *dev = nullptr;
    if (n_obs == 0) return 0;
    CU(h->c_obs.ensure((size_t)n_obs));
    CU(cudaMemcpyAsync(h->c_obs.p, obs, (size_t)n_obs * sizeof(DevObstacle), cudaMemcpyHostToDevice, st));
    *dev = h->c_obs.p;
    return 0;
  }
  int finish() {
    CU(cudaGetLastError());
    for (int i = 0; i < n_out; ++i)
      CU(cudaMemcpyAsync(host_out[i], h->c_out[i].p, out_count[i] * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    h->last_launches = 1;
    return 0;
  }
};
int comp_begin(i2lqr_handle* h, int64_t B, const i2lqr_obstacle* obs, int64_t n_obs) {
  if (!h) return fail("null handle");
  if (B < 1) return fail("batch size must be >= 1");
  if (n_obs != 0 && n_obs != 1 && n_obs != B) return fail("n_obs must be 0, 1 or B");
  if (n_obs != 0 && !obs) return fail("null obstacle array");
  if (h->params.precision != 0) return fail("the component entry points are FP64 only");
  CU(cudaSetDevice(h->device));
  return 0;
}
inline unsigned comp_grid(int64_t B) { return (unsigned)((B + kCompThreads - 1) / kCompThreads); }
}  // namespace

int i2lqr_dynamics_batch(i2lqr_handle* h, int64_t B, const double* x, const double* u, double* x_next) {
  if (comp_begin(h, B, nullptr, 0)) return 1;
  if (!x || !u || !x_next) return fail("null argument");
  CompIO io{h, h->stream};
  const double *dx, *du;
  double* dn;
  if (io.in(x, B * 4, &dx) || io.in(u, B * 2, &du) || io.out(x_next, B * 4, &dn)) return 1;
  comp_dynamics_kernel<<<comp_grid(B), kCompThreads, 0, io.st>>>(h->kp, B, dx, du, dn);
  return io.finish();
}

int i2lqr_linearize_batch(i2lqr_handle* h, int64_t B, const double* v, const double* theta, const double* accel, double* A,
                          double* Bm) {
  if (comp_begin(h, B, nullptr, 0)) return 1;
  if (!theta) return fail("null theta");
  if (A && (!v || !accel)) return fail("get_A_matrix needs velocity and acceleration values");
  const size_t N = h->kp.N;
  CompIO io{h, h->stream};
  const double *dv, *dth, *da;
  double *dA, *dB;
  if (io.in(v ? v : theta, B * N, &dv) || io.in(theta, B * N, &dth) || io.in(accel, B * N, &da)) return 1;
  if (io.out(A, B * 16 * N, &dA) || io.out(Bm, B * 8 * N, &dB)) return 1;
  comp_linearize_kernel<<<comp_grid(B), kCompThreads, 0, io.st>>>(h->kp, B, dv, dth, da, dA, dB);
  return io.finish();
}

int i2lqr_cost_derivatives_batch(i2lqr_handle* h, int64_t B, const double* uvar, const double* dX, const double* xvar,
                                 const i2lqr_obstacle* obs, int64_t n_obs, double* l_u, double* l_uu, double* l_x,
                                 double* l_xx) {
  if (comp_begin(h, B, obs, n_obs)) return 1;
  if (!uvar || !dX || !xvar) return fail("null argument");
  const size_t N = h->kp.N;
  CompIO io{h, h->stream};
  const double *du, *dd, *dx;
  const DevObstacle* dobs;
  double *o0, *o1, *o2, *o3;
  if (io.in(uvar, B * 2 * N, &du) || io.in(dX, B * 4 * (N + 1), &dd) || io.in(xvar, B * 4 * (N + 1), &dx)) return 1;
  if (io.obstacles(obs, n_obs, &dobs)) return 1;
  if (io.out(l_u, B * 2 * N, &o0) || io.out(l_uu, B * 4 * N, &o1) || io.out(l_x, B * 4 * N, &o2) || io.out(l_xx, B * 16 * N, &o3))
    return 1;
  comp_cost_derivatives_kernel<<<comp_grid(B), kCompThreads, 0, io.st>>>(h->kp, B, du, dd, dx, dobs, n_obs, o0, o1, o2, o3);
  return io.finish();
}

int i2lqr_cost_final_batch(i2lqr_handle* h, int64_t B, const double* xvar, const double* x_term, const i2lqr_obstacle* obs,
                           int64_t n_obs, double* l_x, double* l_xx) {
  if (comp_begin(h, B, obs, n_obs)) return 1;
  if (!xvar || !x_term) return fail("null argument");
  const size_t N = h->kp.N;
  CompIO io{h, h->stream};
  const double *dx, *dt;
  const DevObstacle* dobs;
  double *o0, *o1;
  if (io.in(xvar, B * 4 * (N + 1), &dx) || io.in(x_term, B * 4, &dt) || io.obstacles(obs, n_obs, &dobs)) return 1;
  if (io.out(l_x, B * 4, &o0) || io.out(l_xx, B * 16, &o1)) return 1;
  comp_cost_final_kernel<<<comp_grid(B), kCompThreads, 0, io.st>>>(h->kp, B, dx, dt, dobs, n_obs, o0, o1);
  return io.finish();
}

int i2lqr_backward_pass_batch(i2lqr_handle* h, int64_t B, const double* xvar, const double* uvar, const double* x_term,
                              const double* lamb, const i2lqr_obstacle* obs, int64_t n_obs, double* k, double* K) {
  if (comp_begin(h, B, obs, n_obs)) return 1;
  if (!xvar || !uvar || !x_term || !k || !K) return fail("null argument");
  const size_t N = h->kp.N;
  CompIO io{h, h->stream};
  const double *dx, *du, *dt, *dl;
  const DevObstacle* dobs;
  double *o0, *o1;
  if (io.in(xvar, B * 4 * (N + 1), &dx) || io.in(uvar, B * 2 * N, &du) || io.in(x_term, B * 4, &dt) || io.in(lamb, B, &dl)) return 1;
  if (io.obstacles(obs, n_obs, &dobs)) return 1;
  if (io.out(k, B * 2 * N, &o0) || io.out(K, B * 8 * N, &o1)) return 1;
  const long long lanes = (long long)comp_grid(B) * kCompThreads;
  CU(h->c_scratch.ensure((size_t)lanes * 18 * N));
  comp_backward_kernel<<<comp_grid(B), kCompThreads, 0, io.st>>>(h->kp, B, dx, du, dt, dl, dobs, n_obs, h->c_scratch.p, lanes, o0, o1);
  return io.finish();
}

int i2lqr_forward_pass_batch(i2lqr_handle* h, int64_t B, const double* xvar, const double* uvar, const double* x_term,
                             const double* k, const double* K, double* xvar_new, double* uvar_new, double* cost_new) {
  if (comp_begin(h, B, nullptr, 0)) return 1;
  if (!xvar || !uvar || !x_term || !k || !K) return fail("null argument");
  const size_t N = h->kp.N;
  CompIO io{h, h->stream};
  const double *dx, *du, *dt, *dk, *dK;
  double *o0, *o1, *o2;
  if (io.in(xvar, B * 4 * (N + 1), &dx) || io.in(uvar, B * 2 * N, &du) || io.in(x_term, B * 4, &dt) || io.in(k, B * 2 * N, &dk) ||
      io.in(K, B * 8 * N, &dK))
    return 1;
  if (io.out(xvar_new, B * 4 * (N + 1), &o0) || io.out(uvar_new, B * 2 * N, &o1) || io.out(cost_new, B, &o2)) return 1;
  const long long lanes = (long long)comp_grid(B) * kCompThreads;
  CU(h->c_scratch.ensure((size_t)lanes * 18 * N));
  comp_forward_kernel<<<comp_grid(B), kCompThreads, 0, io.st>>>(h->kp, B, dx, du, dt, dk, dK, h->c_scratch.p, lanes, o0, o1, o2);
  return io.finish();
}

// ------------------------------------------------------------------------------------ batched episodes
namespace {
EpisodeState episode_state(i2lqr_handle* h) {
  EpisodeState s{};
  s.x = h->s_x.p;
  s.upred = h->ep_upred.p;
  s.step = h->ep_step.p;
  s.done = h->ep_done.p;
  s.horizon = h->ep_horizon.p;
  s.pb = h->ep_pb.p;
  s.active = h->ep_active.p;
  s.status = h->ep_status.p;
  s.lap_x = h->ep_lapx.p;
  s.lap_u = h->ep_lapu.p;
  s.S = h->params.max_sim_steps;
  s.obs = h->d_obs.p;
  s.obs0 = h->d_obs0.p;
  if (h->extras_in_use > 0) {
    s.obs_extra = h->d_obs_extra.p;
    s.obs_extra0 = h->d_obs_extra0.p;
    s.n_extra = h->d_n_extra.p;
  }
  s.goal = h->ep_goal.p;
  s.n_running = h->ep_nrun.p;
  return s;
}
}  // namespace

int i2lqr_episodes_begin(i2lqr_handle* h, const double* x_init, const double* goal) {
  if (!h) return fail("null handle");
  if (h->E < 1) return fail("handle was created without episodes");
  if (!x_init || !goal) return fail("null argument");
  if (h->params.max_sim_steps < 1) return fail("max_sim_steps must be >= 1");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  if (ensure_step_workspace(h)) return 1;
  const size_t E = h->E, N = h->kp.N, S = h->params.max_sim_steps;
  CU(h->ep_upred.ensure(E * 2 * N));
  CU(h->ep_goal.ensure(4));
  CU(h->ep_lapx.ensure(E * (S + 1) * 4));
  CU(h->ep_lapu.ensure(E * S * 2));
  for (auto* b : {&h->ep_step, &h->ep_done, &h->ep_horizon, &h->ep_pb, &h->ep_active, &h->ep_status, &h->ep_laplen}) CU(b->ensure(E));
  CU(h->ep_nrun.ensure(1));
  if (!h->h_nrun) {
    CU(cudaHostAlloc((void**)&h->h_nrun, 2 * sizeof(int), cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer((void**)&h->h_nrun_dev, h->h_nrun, 0));
    for (auto& e : h->ep_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  CU(cudaMemcpyAsync(h->s_x.p, x_init, E * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->ep_goal.p, goal, 4 * sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(h->ep_step.p, 0, E * sizeof(int), st));
  CU(cudaMemsetAsync(h->ep_done.p, 0, E * sizeof(int), st));
  CU(cudaMemsetAsync(h->ep_pb.p, 0, E * sizeof(int), st));
  CU(cudaMemsetAsync(h->ep_status.p, 0, E * sizeof(int), st));
  std::vector<int> hor(E, (int)N);
  CU(cudaMemcpyAsync(h->ep_horizon.p, hor.data(), E * sizeof(int), cudaMemcpyHostToDevice, st));
  if (spec_prepare(h, st)) return 1;
  CU(cudaStreamSynchronize(st));
  h->ep_ready = true;
  h->ep_lap_index = 0;
  return 0;
}

int i2lqr_episodes_set_noise(i2lqr_handle* h, uint64_t seed) {
  if (!h) return fail("null handle");
  h->ep_noise_seed = seed;
  return 0;
}

int i2lqr_episodes_run_lap(i2lqr_handle* h, int32_t* lap_len, double* lap_x, double* lap_u) {
  if (!h) return fail("null handle");
  if (!h->ep_ready) return fail("call i2lqr_episodes_begin first");
  for (int e = 0; e < h->E; ++e)
    if (h->h_lap_count[e] == 0) return fail("episode %d has no stored lap", e);
  CU(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int E = h->E, S = h->params.max_sim_steps, N = h->kp.N;
  if (sync_obstacles(h)) return 1;
  CU(h->d_obs0.ensure(E));
  CU(cudaMemcpyAsync(h->d_obs0.p, h->d_obs.p, sizeof(DevObstacle) * E, cudaMemcpyDeviceToDevice, st));
  if (h->extras_in_use > 0) {
    CU(h->d_obs_extra0.ensure((size_t)E * kMaxExtraObstacles));
    CU(cudaMemcpyAsync(h->d_obs_extra0.p, h->d_obs_extra.p, sizeof(DevObstacle) * E * kMaxExtraObstacles, cudaMemcpyDeviceToDevice, st));
  }
  EpisodeState es = episode_state(h);
  // row 0 of every lap log = the state the lap starts from
  CU(cudaMemcpy2DAsync(h->ep_lapx.p, (size_t)(S + 1) * 4 * sizeof(double), h->s_x.p, 4 * sizeof(double), 4 * sizeof(double), E,
                       cudaMemcpyDeviceToDevice, st));
  CU(cudaMemsetAsync(h->ep_status.p, 0, (size_t)E * sizeof(int), st));  // status is per lap, not sticky
  // Two words of mapped pinned memory carry [episodes still running, episodes that solve this step] back to the
  // host after every step.  The host never waits for them: both only fall during a lap, so whatever (stale) value
  // it reads is an upper bound -- enough to size the next step's grids and to stop issuing steps once every episode
  // has finished.  No stream synchronisation inside the lap.
  h->h_nrun[0] = E;
  h->h_nrun[1] = E;
  CU(h->ep_list.ensure((size_t)E));
  CU(h->ep_nep.ensure(1));
  CU(cudaMemcpyAsync(h->ep_nrun.p, h->h_nrun, sizeof(int), cudaMemcpyHostToDevice, st));
  if (begin_stats(h, st)) return 1;
  const int tb = 128, gb = (E + tb - 1) / tb;
  volatile int* hw = h->h_nrun;
  int issued = 0;
  static const bool ep_timing = getenv("I2LQR_EP_TIMING") != nullptr;  // developer switch: per-step device times on stderr
  std::vector<cudaEvent_t> tev;
  std::vector<int> tupper;
  auto tmark = [&]() {
    if (!ep_timing) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tev.push_back(e);
  };
  tmark();
  static const int lookahead = getenv("I2LQR_EP_LOOKAHEAD") ? atoi(getenv("I2LQR_EP_LOOKAHEAD")) : 4;  // tuning switch
  for (int i = 0; i < S; ++i) {
    // at most `lookahead` steps in flight: a deeper queue would size many steps from the same stale counts
    if (i >= lookahead) CU(cudaEventSynchronize(h->ep_ev[(i - lookahead) % i2lqr_handle::kEpEvents]));
    if (hw[0] <= 0) break;
    const int n_upper = std::min(E, std::max(1, (int)hw[1]));
    CU(cudaMemsetAsync(h->ep_nep.p, 0, sizeof(int), st));
    episode_pre_kernel<<<gb, tb, 0, st>>>(es, E, N, h->ep_list.p, h->ep_nep.p);
    CU(cudaGetLastError());
    StepScope sc{n_upper, nullptr, h->ep_list.p, h->ep_nep.p};
    if (run_step(h, sc, st, false)) return 1;
    episode_post_kernel<<<gb, tb, 0, st>>>(es, h->kp, h->s_u.p, h->s_upred.p, h->s_shrink.p, E, i, h->params.goal_tol,
                                           h->ep_noise_seed, h->ep_lap_index);
    CU(cudaGetLastError());
    episode_publish_kernel<<<1, 1, 0, st>>>(h->ep_nrun.p, h->ep_nep.p, h->h_nrun_dev);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ep_ev[i % i2lqr_handle::kEpEvents], st));
    h->last_launches += 3;
    ++issued;
    tupper.push_back(n_upper);
    tmark();
  }
  (void)issued;
  if (ep_timing) {
    cudaStreamSynchronize(st);
    fprintf(stderr, "[i2lqr] lap %d: per step (size bound : ms):", h->ep_lap_index);
    for (size_t i = 0; i + 1 < tev.size(); ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, tev[i], tev[i + 1]);
      fprintf(stderr, " %d:%.2f", tupper[i], ms);
    }
    fprintf(stderr, "\n");
    for (auto e : tev) cudaEventDestroy(e);
  }
  episode_finish_kernel<<<(unsigned)(((long long)E * 32 + tb - 1) / tb), tb, 0, st>>>(es, h->set, E, N, h->ep_laplen.p);
  CU(cudaGetLastError());
  h->last_launches += 1;
  if (end_stats(h, st, true)) return 1;
  h->ep_lap_index += 1;
  // host mirrors of the safe-set bookkeeping
  CU(cudaMemcpy(h->h_lap_len.data(), h->set.lap_len, sizeof(int) * (size_t)E * h->L, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(h->h_lap_count.data(), h->set.lap_count, sizeof(int) * (size_t)E, cudaMemcpyDeviceToHost));
  if (lap_len) CU(cudaMemcpy(lap_len, h->ep_laplen.p, sizeof(int) * (size_t)E, cudaMemcpyDeviceToHost));
  if (lap_x) CU(cudaMemcpy(lap_x, h->ep_lapx.p, sizeof(double) * (size_t)E * (S + 1) * 4, cudaMemcpyDeviceToHost));
  if (lap_u) CU(cudaMemcpy(lap_u, h->ep_lapu.p, sizeof(double) * (size_t)E * S * 2, cudaMemcpyDeviceToHost));
  return 0;
}

/* Per-episode status of the last lap: 0 ok, 1 hit max_sim_steps, 2 playback exhausted, 3 lap not storable. */
int i2lqr_episodes_status(i2lqr_handle* h, int32_t* status) {
  if (!h || !status) return fail("null argument");
  if (!h->ep_ready) return fail("call i2lqr_episodes_begin first");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpy(status, h->ep_status.p, sizeof(int) * (size_t)h->E, cudaMemcpyDeviceToHost));
  return 0;
}

int i2lqr_last_stats(i2lqr_handle* h, double* kernel_ms, int64_t* launches, int64_t* ilqr_iterations, int64_t* solves) {
  if (!h) return fail("null handle");
  if (h->stats_lazy) {  // the no-copy step returned before its kernel's tail: read the device time now
    CU(cudaSetDevice(h->device));
    CU(cudaEventSynchronize(h->ev1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    h->stats_lazy = false;
  }
  if (kernel_ms) *kernel_ms = h->last_ms;
  if (launches) *launches = h->last_launches;
  if (ilqr_iterations) *ilqr_iterations = h->last_iters;
  if (solves) *solves = h->last_solves;
  return 0;
}

// Cumulative counters of the speculative one-launch controller step (spec_step.cuh) since the handle was created.
int i2lqr_spec_stats(i2lqr_handle* h, int64_t* counters8) {
  if (!h || !counters8) return fail("null argument");
  for (int i = 0; i < 8; ++i) counters8[i] = 0;
  if (!h->sp_stats.p) return 0;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  unsigned long long v[8];
  CU(cudaMemcpy(v, h->sp_stats.p, sizeof v, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 8; ++i) counters8[i] = (int64_t)v[i];
  return 0;
}

// Developer aid (I2LQR_SPEC_TIMELINE=1): per (outer iteration, chain) of the last one-launch step, 8 words -- globaltimer ns at
// start / K-nearest done / rank-0 candidate done / walk done / released by the winner, ns spent waiting for in-flight entries,
// (in-place solves << 32 | their iterations), 0.  Returns the number of chains the table is laid out for.
int i2lqr_spec_timeline(i2lqr_handle* h, uint64_t* words, int64_t capacity_words, int32_t* c_max) {
  if (!h || !words || !c_max) return fail("null argument");
  *c_max = 0;
  if (!h->sp_timeline.p) return 0;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  const size_t n = std::min<size_t>((size_t)capacity_words, h->sp_timeline.n);  // (chain table first, then 4 words per table entry)
  CU(cudaMemcpy(words, h->sp_timeline.p, n * 8, cudaMemcpyDeviceToHost));
  *c_max = h->sp_cmax;
  return 0;
}

// Refreshes last_stats after an asynchronous *_device call (synchronises the stream the call ran on).
int i2lqr_sync_stats(i2lqr_handle* h, void* cuda_stream) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  return finish_stats(h, st);
}

// FP64 FMA-pipe peak of `device` in TFLOP/s (2 flops per DFMA), best of `reps` timed launches.
int i2lqr_measure_fp64_peak(int device, int reps, double* tflops) {
  if (!tflops) return fail("null argument");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return fail("no such CUDA device");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 15;
  double* out = nullptr;
  CU(cudaMalloc(&out, sizeof(double) * threads * blocks));
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a));
  CU(cudaEventCreate(&b));
  double best = 0;
  for (int r = 0; r < reps + 2; ++r) {
    CU(cudaEventRecord(a));
    dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    CU(cudaEventRecord(b));
    CU(cudaEventSynchronize(b));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, a, b));
    const double tf = 2.0 * 8.0 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
    if (r >= 2 && tf > best) best = tf;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(out);
  *tflops = best;
  return 0;
}

}  // extern "C"
