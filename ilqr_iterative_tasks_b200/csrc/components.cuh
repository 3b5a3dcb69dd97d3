// components.cuh -- the functions the iLQR iteration is made of, as batched entry points of their own.
//
// The solve kernels (kernels.cuh) run these steps fused, with every intermediate in registers / shared memory.
// A caller of the reference can also reach them one by one:
//   kinetic_bicycle, get_A_matrix, get_B_matrix          iterative_ilqr/systems/kinetic_bicycle.py:10-52
//   get_cost_derivation, add_control_constraint,
//   repelling_cost_function, get_cost_final              iterative_ilqr/control/ilqr_helper.py:9-64,83-150
//   backward_pass, forward_pass                          iterative_ilqr/control/iterative_ilqr.py:88-160
// The kernels below expose exactly those, one thread per problem, by calling the SAME device functions the
// solve kernels inline (chain_task_eval of chain_core.cuh for the stage terms, backward_pass / forward_pass of
// ilqr_core.cuh on a global scratch column) and expanding their structured results into the reference's dense
// array shapes (time on the last axis).  They exist for parity at function granularity and for callers that
// drive their own outer loop; the throughput path is the fused kernels.
#pragma once
#include "chain_core.cuh"
#include "ilqr_core.cuh"

namespace i2lqr {

constexpr int kCompThreads = 128;

__device__ __forceinline__ void comp_obstacle(const KParams& p, Cand& c, const DevObstacle* obs, long long n_obs, long long b) {
  load_obstacle(p, c, n_obs == 0 ? nullptr : obs + (n_obs == 1 ? 0 : b));
}

// x (B,4), u (B,2) -> x_next (B,4): the plant / rollout step.
__global__ void comp_dynamics_kernel(const __grid_constant__ KParams p, long long B, const double* x, const double* u, double* xn) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* xb = x + b * 4;
  const double ua = u[b * 2], ud = u[b * 2 + 1];
  double sn, cs;
  sincos_fast(xb[3], &sn, &cs);
  const double sd = xb[2] * p.dt + (ua * p.dt2) / 2.0;  // same expression as init_candidate / forward_pass
  xn[b * 4 + 0] = xb[0] + cs * sd;
  xn[b * 4 + 1] = xb[1] + sn * sd;
  xn[b * 4 + 2] = xb[2] + ua * p.dt;
  xn[b * 4 + 3] = xb[3] + ud * p.dt;
}

// Stage terms of problem b, stage k, from the reference's arrays xvar (4,N+1), uvar (2,N): the linearisation
// point is (v_{k+1}, theta_{k+1}, a_k) (iterative_ilqr.py:92-99), the cost derivatives are taken at x_k, u_k.
__device__ __forceinline__ void comp_task(const KParams& p, const double* xv, const double* uv, int k, TaskIn<double>& t) {
  const int N = p.N, ld = N + 1;
  t.k = k;
#pragma unroll
  for (int i = 0; i < 4; ++i) t.x[i] = xv[i * ld + k];
  const int ku = k < N ? k : N - 1;
  t.ua = uv ? uv[ku] : 0.0;
  t.ud = uv ? uv[N + ku] : 0.0;
  t.v1 = xv[2 * ld + ku + 1];
  sincos_fast(xv[3 * ld + ku + 1], &t.sn, &t.cs);
}

// get_A_matrix(v, theta, a, N, dt) -> A (B,4,4,N);  get_B_matrix(theta, N, dt) -> Bm (B,4,2,N).
// v, theta, accel are (B,N) each, as the reference's vectorised arguments.
__global__ void comp_linearize_kernel(const __grid_constant__ KParams p, long long B, const double* v, const double* theta,
                                      const double* accel, double* A, double* Bm) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int N = p.N;
  Cand c = {};
  for (int k = 0; k < N; ++k) {
    TaskIn<double> t = {};
    t.k = k;
    t.v1 = v[b * N + k];
    t.ua = accel ? accel[b * N + k] : 0.0;
    sincos_fast(theta[b * N + k], &t.sn, &t.cs);
    double o[kTermWidth];
    chain_task_eval<false>(p, c, t, o);
    if (A) {
      double* a = A + b * 16 * N + k;
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i * N] = (i % 5 == 0) ? 1.0 : 0.0;
      a[2 * N] = o[0];   // cos(theta) dt
      a[3 * N] = o[1];   // -(v dt + a dt^2 / 2) sin(theta)
      a[6 * N] = o[2];   // sin(theta) dt
      a[7 * N] = o[3];   // (v dt + a dt^2 / 2) cos(theta)
    }
    if (Bm) {
      double* m = Bm + b * 8 * N + k;
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i * N] = 0.0;
      m[0 * N] = o[4];   // dt^2 cos(theta) / 2
      m[2 * N] = o[5];   // dt^2 sin(theta) / 2
      m[4 * N] = p.dt;
      m[7 * N] = p.dt;
    }
  }
}

// get_cost_derivation(ctrl_U, dX, ..., xvar, obstacle, ...) -> l_u (B,2,N), l_uu (B,2,2,N), l_x (B,4,N), l_xx (B,4,4,N).
// dX is an argument of its own in the reference (ilqr() keeps it equal to xvar - xtarget): l_x = 2 Q dX[:,k] + barrier.
__global__ void comp_cost_derivatives_kernel(const __grid_constant__ KParams p, long long B, const double* uvar, const double* dX,
                                             const double* xvar, const DevObstacle* obs, long long n_obs, double* l_u,
                                             double* l_uu, double* l_x, double* l_xx) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int N = p.N, ld = N + 1;
  Cand c = {};
  comp_obstacle(p, c, obs, n_obs, b);
  const double* xv = xvar + b * 4 * ld;
  const double* uv = uvar + b * 2 * N;
  const double* dx = dX + b * 4 * ld;
  for (int k = 0; k < N; ++k) {
    TaskIn<double> t;
    comp_task(p, xv, uv, k, t);
    double o[kTermWidth];
    chain_task_eval<false>(p, c, t, o);  // barrier terms only; the Q / R terms are added below from dX
    const double d0 = dx[k], d1 = dx[ld + k], d2 = dx[2 * ld + k], d3 = dx[3 * ld + k];
    const double ua = t.ua, ud = t.ud;
    if (l_u) {
      l_u[b * 2 * N + k] = (p.R2[0] * ua + p.R2[1] * ud) + o[6];
      l_u[b * 2 * N + N + k] = (p.R2[1] * ua + p.R2[2] * ud) + o[7];
    }
    if (l_uu) {
      double* m = l_uu + b * 4 * N + k;
      m[0] = o[8] + p.R2[0];
      m[N] = p.R2[1];
      m[2 * N] = p.R2[1];
      m[3 * N] = o[9] + p.R2[2];
    }
    if (l_x) {
      double* g = l_x + b * 4 * N + k;
      g[0] = (p.Q2[0] * d0 + p.Q2[1] * d1 + p.Q2[2] * d2 + p.Q2[3] * d3) + o[10];
      g[N] = (p.Q2[1] * d0 + p.Q2[4] * d1 + p.Q2[5] * d2 + p.Q2[6] * d3) + o[11];
      g[2 * N] = p.Q2[2] * d0 + p.Q2[5] * d1 + p.Q2[7] * d2 + p.Q2[8] * d3;
      g[3 * N] = p.Q2[3] * d0 + p.Q2[6] * d1 + p.Q2[8] * d2 + p.Q2[9] * d3;
    }
    if (l_xx) {
      double* h = l_xx + b * 16 * N + k;
      const int tri[16] = {0, 1, 2, 3, 1, 4, 5, 6, 2, 5, 7, 8, 3, 6, 8, 9};
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i * N] = p.Q2[tri[i]];
      h[0] = p.Q2[0] + o[12];
      h[N] = p.Q2[1] + o[13];
      h[4 * N] = p.Q2[1] + o[13];
      h[5 * N] = p.Q2[4] + o[14];
    }
  }
}

// get_cost_final(x, x_terminal, Qterminal, obstacle, ilqr_param) -> l_x (B,4), l_xx (B,4,4)
__global__ void comp_cost_final_kernel(const __grid_constant__ KParams p, long long B, const double* xvar, const double* x_term,
                                       const DevObstacle* obs, long long n_obs, double* Vx, double* Vxx) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int N = p.N, ld = N + 1;
  Cand c = {};
  comp_obstacle(p, c, obs, n_obs, b);
#pragma unroll
  for (int i = 0; i < 4; ++i) c.xt[i] = x_term[b * 4 + i];
  TaskIn<double> t;
  comp_task(p, xvar + b * 4 * ld, nullptr, N, t);
  double o[kTermWidth];
  chain_task_eval<false>(p, c, t, o);
  if (Vx) {
    Vx[b * 4 + 0] = o[15];
    Vx[b * 4 + 1] = o[16];
    Vx[b * 4 + 2] = o[17];
    Vx[b * 4 + 3] = o[18];
  }
  if (Vxx) {
    double* h = Vxx + b * 16;
    const int tri[16] = {0, 1, 2, 3, 1, 4, 5, 6, 2, 5, 7, 8, 3, 6, 8, 9};
#pragma unroll
    for (int i = 0; i < 16; ++i) h[i] = p.Qf2[tri[i]];
    h[0] = p.Qf2[0] + o[12];
    h[1] = p.Qf2[1] + o[13];
    h[4] = p.Qf2[1] + o[13];
    h[5] = p.Qf2[4] + o[14];
  }
}

// Loads a problem's trajectory into the lane's scratch column in the solve kernels' slot format.
__device__ __forceinline__ void comp_load_trajectory(const KParams& p, Cand& c, const GlobalScratch<double>& s, const double* xv,
                                                     const double* uv) {
  const int N = p.N, ld = N + 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) c.x0[i] = xv[i * ld];
  sincos_fast(c.x0[3], &c.s0, &c.c0);
  for (int k = 0; k < N; ++k) {
    s.stc(I2_CUR(k) + 0, uv[k]);
    s.stc(I2_CUR(k) + 1, uv[N + k]);
#pragma unroll
    for (int i = 0; i < 4; ++i) s.stc(I2_CUR(k) + 2 + i, xv[i * ld + k + 1]);
    double sn, cs;
    sincos_fast(xv[3 * ld + k + 1], &sn, &cs);
    s.stc(I2_CUR(k) + 6, cs);
    s.stc(I2_CUR(k) + 7, sn);
  }
}

// backward_pass(xvar, uvar, x_terminal, dX, lamb, ...) -> k (B,2,N), K (B,2,4,N).  dX is taken as xvar - xtarget,
// which is what ilqr() passes (iterative_ilqr.py:41-42).
__global__ void comp_backward_kernel(const __grid_constant__ KParams p, long long B, const double* xvar, const double* uvar,
                                     const double* x_term, const double* lamb, const DevObstacle* obs, long long n_obs,
                                     double* gscratch, long long gstride, double* k_out, double* K_out) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int N = p.N, ld = N + 1;
  Cand c = {};
  comp_obstacle(p, c, obs, n_obs, b);
#pragma unroll
  for (int i = 0; i < 4; ++i) c.xt[i] = x_term[b * 4 + i];
  c.lamb = lamb ? lamb[b] : p.lamb0;
  GlobalScratch<double> s;
  s.bind(nullptr, gscratch + b, gstride, N);
  comp_load_trajectory(p, c, s, xvar + b * 4 * ld, uvar + b * 2 * N);
  backward_pass<true, 0>(p, c, s);
  for (int k = 0; k < N; ++k) {  // the solve kernels keep the gains sign-flipped (ilqr_core.cuh backward_pass)
    k_out[b * 2 * N + k] = -s.ldg(k * 10 + 0);
    k_out[b * 2 * N + N + k] = -s.ldg(k * 10 + 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) K_out[b * 8 * N + i * N + k] = -s.ldg(k * 10 + 2 + i);
  }
}

// forward_pass(xvar, uvar, x_terminal, ..., matrix_k, matrix_K, ...) -> xvar_new (B,4,N+1), uvar_new (B,2,N), cost_new (B)
__global__ void comp_forward_kernel(const __grid_constant__ KParams p, long long B, const double* xvar, const double* uvar,
                                    const double* x_term, const double* k_in, const double* K_in, double* gscratch,
                                    long long gstride, double* xvar_new, double* uvar_new, double* cost_new) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int N = p.N, ld = N + 1;
  Cand c = {};
#pragma unroll
  for (int i = 0; i < 4; ++i) c.xt[i] = x_term[b * 4 + i];
  GlobalScratch<double> s;
  s.bind(nullptr, gscratch + b, gstride, N);
  comp_load_trajectory(p, c, s, xvar + b * 4 * ld, uvar + b * 2 * N);
  for (int k = 0; k < N; ++k) {
    s.stg(k * 10 + 0, -k_in[b * 2 * N + k]);
    s.stg(k * 10 + 1, -k_in[b * 2 * N + N + k]);
#pragma unroll
    for (int i = 0; i < 8; ++i) s.stg(k * 10 + 2 + i, -K_in[b * 8 * N + i * N + k]);
  }
  double Jt;
  const double Jn = forward_pass<true, 0>(p, c, s, Jt);
  if (cost_new) cost_new[b] = Jn;
  if (xvar_new) {
    double* xo = xvar_new + b * 4 * ld;
#pragma unroll
    for (int i = 0; i < 4; ++i) xo[i * ld] = c.x0[i];
    for (int k = 0; k < N; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) xo[i * ld + k + 1] = s.ldg(k * 10 + 2 + i);
  }
  if (uvar_new)
    for (int k = 0; k < N; ++k) {
      uvar_new[b * 2 * N + k] = s.ldg(k * 10 + 0);
      uvar_new[b * 2 * N + N + k] = s.ldg(k * 10 + 1);
    }
}

}  // namespace i2lqr
