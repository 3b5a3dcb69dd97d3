// Developer tool (not product, not shipped in any library): compiles csrc/ilqr_core.cuh for the HOST so that
// algorithmic rewrites of the device code (closed-form inverse, fused exponentials, ...) can be checked
// against the golden vectors here, where there is no GPU.  Host libm and host FMA contraction differ from the
// device by ulps, so this only screens for discrete-path regressions before a GPU run.
//   g++ -O2 -mfma -ffp-contract=fast -std=c++17 -shared -fPIC -o /tmp/libhostcore.so tools/host_core_check.cpp
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#define __device__
#define __host__
#define __forceinline__ inline
#define I2LQR_HOST_BUILD 1
#include "../ilqr_iterative_tasks_b200/csrc/ilqr_core.cuh"
#include "../ilqr_iterative_tasks_b200/csrc/chain_core.cuh"
#include "../ilqr_iterative_tasks_b200/csrc/kparams.h"
using namespace i2lqr;

extern "C" const char* host_fill_kparams(const i2lqr_params* p, KParams* k) { return fill_kparams(*p, k); }
extern "C" int host_kparams_size() { return (int)sizeof(KParams); }

extern "C" int host_solve(const KParams* kp, int64_t B, const double* x0, const double* xt, const double* lamb_in,
                          const i2lqr_obstacle* obs, int64_t n_obs, const double* u_init, double* uvar, double* xvar,
                          double* lamb_out, int32_t* iters) {
  const KParams& p = *kp;
  std::vector<double> mem(scratch_doubles(p.N));
  GlobalScratch<double> s;
  s.bind(nullptr, mem.data(), 1, p.N);
  for (int64_t b = 0; b < B; ++b) {
    Cand c;
    for (int i = 0; i < 4; ++i) { c.x0[i] = x0[b * 4 + i]; c.xt[i] = xt[b * 4 + i]; }
    sincos_fast(c.x0[3], &c.s0, &c.c0);
    c.lamb = lamb_in ? lamb_in[b] : p.lamb0;
    load_obstacle(p, c, n_obs == 0 ? nullptr : reinterpret_cast<const DevObstacle*>(n_obs == 1 ? obs : obs + b));
    const double* u0 = u_init ? u_init + b * 2 * p.N : nullptr;
    if (p.has_qr) {
      init_candidate<true, 0>(p, c, s, u0);
      while (!ilqr_iteration<true, 0>(p, c, s)) {}
    } else {
      init_candidate<false, 0>(p, c, s, u0);
      while (!ilqr_iteration<false, 0>(p, c, s)) {}
    }
    store_candidate<0>(p, c, s, xvar + b * 4 * (p.N + 1), uvar + b * 2 * p.N);
    lamb_out[b] = c.lamb;
    iters[b] = c.iters;
  }
  return 0;
}

extern "C" void host_exp_fast(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = exp_fast(x[i]); }
extern "C" void host_sincos_fast(const double* x, double* s, double* c, int64_t n) { for (int64_t i = 0; i < n; ++i) sincos_fast(x[i], s + i, c + i); }

// FP32 build of the same core (inputs / outputs double)
extern "C" int host_solve_f32(const KParams* kp, int64_t B, const double* x0, const double* xt, const double* lamb_in,
                              const i2lqr_obstacle* obs, int64_t n_obs, double* uvar, double* xvar, double* lamb_out,
                              int32_t* iters) {
  const KParamsT<float> p = to_float(*kp);
  std::vector<float> mem(scratch_doubles(p.N));
  GlobalScratch<float> s;
  s.bind(nullptr, mem.data(), 1, p.N);
  for (int64_t b = 0; b < B; ++b) {
    CandT<float> c = {};
    for (int i = 0; i < 4; ++i) { c.x0[i] = (float)x0[b * 4 + i]; c.xt[i] = (float)xt[b * 4 + i]; }
    sincos_fast(c.x0[3], &c.s0, &c.c0);
    c.lamb = lamb_in ? (float)lamb_in[b] : p.lamb0;
    load_obstacle(p, c, n_obs == 0 ? nullptr : reinterpret_cast<const DevObstacle*>(n_obs == 1 ? obs : obs + b));
    if (p.has_qr) {
      init_candidate<true, 0>(p, c, s, nullptr);
      while (!ilqr_iteration<true, 0>(p, c, s)) {}
    } else {
      init_candidate<false, 0>(p, c, s, nullptr);
      while (!ilqr_iteration<false, 0>(p, c, s)) {}
    }
    store_candidate<0>(p, c, s, xvar + b * 4 * (p.N + 1), uvar + b * 2 * p.N);
    lamb_out[b] = c.lamb;
    iters[b] = c.iters;
  }
  return 0;
}

// Host emulation of one lane group of chain_lat_kernel (kernels.cuh): the phases between two __syncwarp()s run
// lane after lane.  One chain: candidates xt[0..n) solved in order from x0, lambda carried.
template <bool kQR>
static void chain_emulate(const KParams& p, int rungs, int n, const double* x0, const double* xt, const DevObstacle* obs,
                          double* uvar, double* xvar, double* lamb_out, int32_t* iters) {
  constexpr int S = 32;
  const int N = p.N;
  std::vector<double> smem((size_t)chain_lane_elems(N, rungs) * S, 0.0);
  std::vector<ChainLane<double, S>> v(rungs);
  std::vector<Cand> c(rungs);
  for (int r = 0; r < rungs; ++r) {
    v[r].col = smem.data() + r;
    v[r].grp = smem.data();
    v[r].N = N;
    v[r].rungs = rungs;
    for (int i = 0; i < 4; ++i) c[r].x0[i] = x0[i];
    sincos_fast(c[r].x0[3], &c[r].s0, &c[r].c0);
    c[r].lamb = p.lamb0;
    load_obstacle(p, c[r], obs);
  }
  std::vector<int> wb(rungs);
  for (int j = 0; j < n; ++j) {
    for (int r = 0; r < rungs; ++r) {
      for (int i = 0; i < 4; ++i) c[r].xt[i] = xt[j * 4 + i];
      c[r].J = chain_init<kQR, S>(p, c[r], r == 0 ? v[r].trial(0) : nullptr);
      c[r].iters = 0;
      wb[r] = r == 0 ? 1 : 0;
    }
    const double* cur = v[0].trial_of(0, 0);
    bool dirty = true;
    for (;;) {
      if (dirty)
        for (int r = 0; r < rungs; ++r) chain_terms<kQR, S>(p, c[r], v[r], cur, r);
      std::vector<double> Jn(rungs), Jt(rungs);
      for (int r = 0; r < rungs; ++r) {
        double lamb_r = c[r].lamb;
        for (int q = 0; q < r; ++q) lamb_r *= p.lamb_factor;
        chain_riccati<kQR, S>(p, v[r], lamb_r);
        Jn[r] = chain_forward<kQR, S>(p, c[r], cur, v[r].gains(), v[r].trial(wb[r]), Jt[r]);
      }
      int winner = -1, consumed = 0;
      bool fin = false;
      double lamb = c[0].lamb;
      for (int r = 0; r < rungs; ++r) {
        if (winner >= 0 || fin) break;
        if (c[0].iters + consumed >= p.max_iter) { fin = true; break; }
        ++consumed;
        if (Jn[r] < c[0].J) { winner = r; lamb /= p.lamb_factor; }
        else { lamb *= p.lamb_factor; if (lamb > p.max_lamb) fin = true; }
      }
      dirty = winner >= 0;
      double Jnew = c[0].J;
      if (winner >= 0) {
        cur = v[0].trial_of(winner, wb[winner]);
        wb[winner] ^= 1;
        if (std::fabs((Jn[winner] - c[0].J) / c[0].J) < p.eps) fin = true;
        Jnew = kQR ? Jt[winner] : Jn[winner];
      }
      for (int r = 0; r < rungs; ++r) { c[r].J = Jnew; c[r].lamb = lamb; c[r].iters += consumed; }
      if (c[0].iters >= p.max_iter) fin = true;
      if (fin) break;
    }
    double* xv = xvar + (size_t)j * 4 * (N + 1);
    double* uv = uvar + (size_t)j * 2 * N;
    for (int i = 0; i < 4; ++i) xv[i * (N + 1)] = x0[i];
    for (int k = 0; k < N; ++k) {
      const double* s = cur + (size_t)(8 * k) * S;
      uv[k] = s[0];
      uv[N + k] = s[S];
      for (int i = 0; i < 4; ++i) xv[i * (N + 1) + k + 1] = s[(2 + i) * S];
    }
    lamb_out[j] = c[0].lamb;
    iters[j] = c[0].iters;
  }
}

extern "C" int host_chain_solve(const KParams* kp, int rungs, int n, const double* x0, const double* xt, const i2lqr_obstacle* obs,
                                double* uvar, double* xvar, double* lamb_out, int32_t* iters) {
  const DevObstacle* o = reinterpret_cast<const DevObstacle*>(obs);
  if (kp->has_qr) chain_emulate<true>(*kp, rungs, n, x0, xt, o, uvar, xvar, lamb_out, iters);
  else chain_emulate<false>(*kp, rungs, n, x0, xt, o, uvar, xvar, lamb_out, iters);
  return 0;
}
