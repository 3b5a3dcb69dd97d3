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
