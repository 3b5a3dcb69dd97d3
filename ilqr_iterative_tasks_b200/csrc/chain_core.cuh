// chain_core.cuh -- latency-oriented iLQR iteration for the controller-step form (lambda chains).
//
// A controller step is a handful of sequential chains (utils/base.py:391-426), so its kernel runs one warp
// per SM sub-partition and is bound by dependent-instruction latency, not by the FP64 pipe: the throughput
// core (ilqr_core.cuh) issues one instruction every 3.2 cycles there (profiles/r01_chain_kernel_rungs4).
// This core computes the SAME arithmetic, reorganised for a group of kRungs lanes that own one chain
// (kernels.cuh, chain_kernel):
//   * the group shares ONE copy of the current trajectory; a rung's trial trajectory becomes the current one
//     by switching a pointer, nothing is copied on accept;
//   * everything in the backward pass that does not depend on lambda or on the value function -- input and
//     obstacle barrier exponentials, the linearisation (ilqr_helper.py:9-56,83-150, kinetic_bicycle.py:30-52)
//     -- is one "task" per stage (plus one for the terminal cost), the tasks are dealt to the group's lanes
//     and computed once per accepted trajectory instead of once per rung and per rejected iteration;
//   * the Riccati recursion (iterative_ilqr.py:111-129) then only reads those terms, prefetching the next
//     stage's while the current one's dependent chain (Quu -> inverse -> gains -> value function) drains;
//   * the forward rollout (iterative_ilqr.py:133-160) is skewed by one stage: sin/cos of heading k (needed for
//     the position update of stage k only) is evaluated side by side with the feedback law of stage k, so the
//     two long dependent chains of a stage overlap in one basic block.
// Shared-memory layout: element e of lane l at smem[e * S + l] (S = 32), per lane
//   [0, 10N) gains | [10N, 18N) trial buffer 0 | [18N, 26N) trial buffer 1 | [26N, 26N + slots*kTermWidth) terms
// where the terms of task k live in lane (k % kRungs) of the group, slot k / kRungs.
#pragma once
#include "ilqr_core.cuh"

namespace i2lqr {

constexpr int kTermWidth = 20;
// term slots: 0 al 1 be 2 ga 3 de 4 b0 5 b1 | 6 lu0 7 lu1 8 luu00 9 luu11 | 10 lx0 (g0) 11 lx1 (g1) 12 l00 13 l01 14 l11
//             | stage task: 15 lx2 16 lx3 (kQR) | terminal task (k == N): 15..18 Vx, 12..14 barrier Hessian

__host__ __device__ __forceinline__ int chain_term_slots(int N, int rungs) { return (N + 1 + rungs - 1) / rungs; }
__host__ __device__ __forceinline__ int chain_lane_elems(int N, int rungs) { return 26 * N + chain_term_slots(N, rungs) * kTermWidth; }

// sin / cos without the range branch in the main path: the branch-free evaluation runs unconditionally (so that
// it shares a basic block with independent work) and the library path patches the rare huge argument afterwards.
template <typename R>
__device__ __forceinline__ void sincos_inline(R x, R* sn, R* cs) {
  sincos_fast(x, sn, cs);
}
template <>
__device__ __forceinline__ void sincos_inline<double>(double x, double* sn, double* cs) {
  const double magic = 6755399441055744.0;
  const double t = fma_(x, 6.36619772367581382433e-01, magic);
  const int k = (int)(double_to_bits(t) & 0xffffffffLL);
  const double kf = t - magic;
  double r = fma_(kf, -1.57079632673412561417e+00, x);
  r = fma_(kf, -6.07710050630396597660e-11, r);
  r = fma_(kf, -2.02226624879595063154e-21, r);
  const double z = r * r;
  double ps = kSinC[0], pc = kCosC[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) {
    ps = fma_(ps, z, kSinC[i]);
    pc = fma_(pc, z, kCosC[i]);
  }
  const double s = fma_(ps * z, r, r);
  const double c = fma_(pc * z, z, fma_(-0.5, z, 1.0));
  const bool swap = (k & 1) != 0;
  const double ss = swap ? c : s, cc = swap ? s : c;
  *sn = (k & 2) ? -ss : ss;
  *cs = ((k + 1) & 2) ? -cc : cc;
}
template <typename R>
__device__ __forceinline__ void sincos_fixup(R x, R* sn, R* cs) {
  if (!(fabs(x) <= R(1.0e5))) sincos_fast(x, sn, cs);  // takes the library path inside
}

// View of the group's shared memory from one lane.
template <typename R, int S>
struct ChainLane {
  R* col;        // this lane's column (element e at col[e * S])
  R* grp;        // column of the group's first lane
  int N, rungs;
  __device__ __forceinline__ R* gains() const { return col; }
  __device__ __forceinline__ R* trial(int buf) const { return col + (size_t)(10 * N + 8 * N * buf) * S; }
  // trial buffer `buf` of the group's lane `rung`
  __device__ __forceinline__ const R* trial_of(int rung, int buf) const { return grp + rung + (size_t)(10 * N + 8 * N * buf) * S; }
  __device__ __forceinline__ R* terms_w(int k) const {  // own slot for task k (k % rungs must be this lane's rung)
    return col + (size_t)(26 * N + (k / rungs) * kTermWidth) * S;
  }
  __device__ __forceinline__ const R* terms_r(int k) const {
    return grp + (k % rungs) + (size_t)(26 * N + (k / rungs) * kTermWidth) * S;
  }
};

// Candidate start (utils/base.py:405-410 and the first rollout of iterative_ilqr.py:32-48) with zero inputs: heading
// and speed stay put.  Returns the cost of the initial trajectory; `dst` (may be null) receives it.
template <bool kQR, int S, typename R>
__device__ __forceinline__ R chain_init(const KParamsT<R>& p, const CandT<R>& c, R* dst) {
  R x[4] = {c.x0[0], c.x0[1], c.x0[2], c.x0[3]};
  R J = R(0.0);
  for (int k = 0; k < p.N; ++k) {
    const R ua = R(0.0), ud = R(0.0);
    if (kQR) {
      const R l_state = quad4(p.Q, x[0] - p.xtarget[0], x[1] - p.xtarget[1], x[2] - p.xtarget[2], x[3] - p.xtarget[3]);
      J = J + l_state + quad2(p.Rm, ua, ud);
    }
    const R sd = x[2] * p.dt + ua * p.hdt2;
    x[0] = x[0] + c.c0 * sd;
    x[1] = x[1] + c.s0 * sd;
    x[2] = x[2] + ua * p.dt;
    x[3] = x[3] + ud * p.dt;
    if (dst != nullptr) {
      R* d = dst + (size_t)(8 * k) * S;
      d[0] = ua;
      d[S] = ud;
      d[2 * S] = x[0];
      d[3 * S] = x[1];
      d[4 * S] = x[2];
      d[5 * S] = x[3];
      d[6 * S] = c.c0;
      d[7 * S] = c.s0;
    }
  }
  return J + terminal_cost(p, c, x);
}

// One lambda-independent task: k < N = stage k (linearisation at (v_{k+1}, theta_{k+1}, a_k), input barriers, obstacle
// barrier at x_k, state-cost gradient), k == N = terminal cost (get_cost_final).  `cur` is the current trajectory.
template <typename R>
struct TaskIn {
  R x[4], ua, ud, v1, cs, sn;
  int k;
};
template <int S, typename R>
__device__ __forceinline__ void chain_task_load(const KParamsT<R>& p, const CandT<R>& c, const R* cur, int k, TaskIn<R>& t) {
  const int N = p.N;
  t.k = k;
  if (k == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) t.x[i] = c.x0[i];
  } else {
    const R* s = cur + (size_t)(8 * (k - 1)) * S;
#pragma unroll
    for (int i = 0; i < 4; ++i) t.x[i] = s[(2 + i) * S];
  }
  const R* s = cur + (size_t)(8 * (k < N ? k : N - 1)) * S;
  t.ua = s[0];
  t.ud = s[S];
  t.v1 = s[4 * S];
  t.cs = s[6 * S];
  t.sn = s[7 * S];
}
template <bool kQR, typename R>
__device__ __forceinline__ void chain_task_eval(const KParamsT<R>& p, const CandT<R>& c, const TaskIn<R>& t, R (&o)[kTermWidth]) {
  const R dt = p.dt;
  const bool terminal = t.k == p.N;
  const R ua = t.ua, ud = t.ud, cs = t.cs, sn = t.sn;
  const R sp = t.v1 * dt + ua * p.hdt2;
  o[0] = cs * dt;
  o[1] = -sp * sn;
  o[2] = sn * dt;
  o[3] = sp * cs;
  o[4] = p.hdt2 * cs;
  o[5] = p.hdt2 * sn;
  R lu0, lu1, luu00, luu11;
  {
    const R e1 = exp_fast_bounded(p.cq2 * (ua - p.a_max)), e3 = exp_fast_bounded(p.cq2 * (ud - p.d_lim));
    const R e2 = p.ca_lo * rcp_fast(e1), e4 = p.cd_lo * rcp_fast(e3);
    lu0 = p.cg * e1 - p.cg * e2;
    lu1 = p.cg * e3 - p.cg * e4;
    luu00 = p.ch * e1 + p.ch * e2;
    luu11 = p.ch * e3 + p.ch * e4;
  }
  R g0, g1, l00, l01, l11;
  obstacle_barrier(p, c, t.x[0], t.x[1], terminal ? p.N_param : t.k, g0, g1, l00, l01, l11);
  if (c.n_ex > 0) obstacle_barrier_extra(p, c, t.x[0], t.x[1], terminal ? p.N_param : t.k, g0, g1, l00, l01, l11);
  R lx2 = R(0.0), lx3 = R(0.0);
  if (kQR) {
    const R d0 = t.x[0] - p.xtarget[0], d1 = t.x[1] - p.xtarget[1], d2 = t.x[2] - p.xtarget[2], d3 = t.x[3] - p.xtarget[3];
    const R lx0 = (p.Q2[0] * d0 + p.Q2[1] * d1 + p.Q2[2] * d2 + p.Q2[3] * d3) + g0;
    const R lx1 = (p.Q2[1] * d0 + p.Q2[4] * d1 + p.Q2[5] * d2 + p.Q2[6] * d3) + g1;
    lx2 = p.Q2[2] * d0 + p.Q2[5] * d1 + p.Q2[7] * d2 + p.Q2[8] * d3;
    lx3 = p.Q2[3] * d0 + p.Q2[6] * d1 + p.Q2[8] * d2 + p.Q2[9] * d3;
    lu0 = (p.R2[0] * ua + p.R2[1] * ud) + lu0;
    lu1 = (p.R2[1] * ua + p.R2[2] * ud) + lu1;
    luu00 += p.R2[0];
    luu11 += p.R2[2];
    o[10] = terminal ? g0 : lx0;
    o[11] = terminal ? g1 : lx1;
  } else {
    o[10] = g0;
    o[11] = g1;
  }
  o[6] = lu0;
  o[7] = lu1;
  o[8] = luu00;
  o[9] = luu11;
  o[12] = l00;
  o[13] = l01;
  o[14] = l11;
  // terminal cost gradient 2 Qf (x_N - x_t) + barrier gradient (ilqr_helper.py:106-150)
  R Vx0, Vx1, Vx2, Vx3;
  {
    const R e0 = t.x[0] - c.xt[0], e1 = t.x[1] - c.xt[1], e2 = t.x[2] - c.xt[2], e3 = t.x[3] - c.xt[3];
    Vx0 = p.Qf2[0] * e0 + p.Qf2[1] * e1 + p.Qf2[2] * e2 + p.Qf2[3] * e3;
    Vx1 = p.Qf2[1] * e0 + p.Qf2[4] * e1 + p.Qf2[5] * e2 + p.Qf2[6] * e3;
    Vx2 = p.Qf2[2] * e0 + p.Qf2[5] * e1 + p.Qf2[7] * e2 + p.Qf2[8] * e3;
    Vx3 = p.Qf2[3] * e0 + p.Qf2[6] * e1 + p.Qf2[8] * e2 + p.Qf2[9] * e3;
    Vx0 += g0;
    Vx1 += g1;
  }
  o[15] = terminal ? Vx0 : lx2;
  o[16] = terminal ? Vx1 : lx3;
  o[17] = Vx2;
  o[18] = Vx3;
  o[19] = R(0.0);
}

// The tasks of one lane, two at a time so that their exponentials interleave.  When the group is smaller than the task
// list (some lane has a second task) EVERY lane runs the two-task body -- a lane without a second task recomputes its
// first one and drops the result -- so that the lanes of a warp do not split into a two-task and a one-task pass.
template <bool kQR, int S, typename R>
__device__ __forceinline__ void chain_terms(const KParamsT<R>& p, const CandT<R>& c, const ChainLane<R, S>& v, const R* cur, int rung) {
  const int N = p.N, n_tasks = N + 1;
  const bool pairs = n_tasks > v.rungs;  // uniform over the warp
  for (int k0 = rung; k0 < n_tasks; k0 += 2 * v.rungs) {
    const int k1 = k0 + v.rungs;
    const bool two = k1 < n_tasks;
    TaskIn<R> a;
    chain_task_load<S>(p, c, cur, k0, a);
    R oa[kTermWidth];
    if (pairs) {  // two tasks side by side
      TaskIn<R> b;
      chain_task_load<S>(p, c, cur, two ? k1 : k0, b);
      R ob[kTermWidth];
      chain_task_eval<kQR>(p, c, a, oa);
      chain_task_eval<kQR>(p, c, b, ob);
      if (two) {
        R* db = v.terms_w(k1);
#pragma unroll
        for (int i = 0; i < kTermWidth - 1; ++i) db[i * S] = ob[i];
      }
    } else {
      chain_task_eval<kQR>(p, c, a, oa);
    }
    R* da = v.terms_w(k0);
#pragma unroll
    for (int i = 0; i < kTermWidth - 1; ++i) da[i * S] = oa[i];
  }
}

// Riccati recursion from the shared terms with this rung's lambda; gains go to the lane's own column.
template <bool kQR, int S, typename R>
__device__ __forceinline__ void chain_riccati(const KParamsT<R>& p, const ChainLane<R, S>& v, R lamb) {
  const int N = p.N;
  const R dt = p.dt;
  R Vx0, Vx1, Vx2, Vx3;
  R V00 = p.Qf2[0], V01 = p.Qf2[1], V02 = p.Qf2[2], V03 = p.Qf2[3], V11 = p.Qf2[4], V12 = p.Qf2[5], V13 = p.Qf2[6],
    V22 = p.Qf2[7], V23 = p.Qf2[8], V33 = p.Qf2[9];
  {
    const R* t = v.terms_r(N);
    Vx0 = t[15 * S];
    Vx1 = t[16 * S];
    Vx2 = t[17 * S];
    Vx3 = t[18 * S];
    V00 += t[12 * S];
    V01 += t[13 * S];
    V11 += t[14 * S];
  }
  constexpr int kT = kQR ? 17 : 15;
  R tn[kT];
  {
    const R* t = v.terms_r(N - 1);
#pragma unroll
    for (int i = 0; i < kT; ++i) tn[i] = t[i * S];
  }
  R* gains = v.gains();
  for (int k = N - 1; k >= 0; --k) {
    R t[kT];
#pragma unroll
    for (int i = 0; i < kT; ++i) t[i] = tn[i];
    {
      const R* nx = v.terms_r(k > 0 ? k - 1 : 0);  // next stage's terms, consumed one trip later
#pragma unroll
      for (int i = 0; i < kT; ++i) tn[i] = nx[i * S];
    }
    const R al = t[0], be = t[1], ga = t[2], de = t[3], b0 = t[4], b1 = t[5];
    const R lu0 = t[6], lu1 = t[7], luu00 = t[8], luu11 = t[9];
    const R l00 = t[12], l01 = t[13], l11 = t[14];
    // ---- Q-function (iterative_ilqr.py:112-116): the same nested fma chains as ilqr_core.cuh backward_pass
    const R G0 = fma(dt, V02, fma(b1, V01, b0 * V00)), G1 = fma(dt, V12, fma(b1, V11, b0 * V01));
    const R G2 = fma(dt, V22, fma(b1, V12, b0 * V02)), G3 = fma(dt, V23, fma(b1, V13, b0 * V03));
    const R H0 = dt * V03, H1 = dt * V13, H2 = dt * V23, H3 = dt * V33;
    const R X02 = fma(al, G0, fma(ga, G1, G2)), X03 = fma(be, G0, fma(de, G1, G3));
    const R X12 = fma(al, H0, fma(ga, H1, H2)), X13 = fma(be, H0, fma(de, H1, H3));
    const R W02 = fma(al, V00, fma(ga, V01, V02)), W12 = fma(al, V01, fma(ga, V11, V12));
    const R W22 = fma(al, V02, fma(ga, V12, V22)), W03 = fma(be, V00, fma(de, V01, V03));
    const R W13 = fma(be, V01, fma(de, V11, V13)), W23 = fma(be, V02, fma(de, V12, V23));
    const R W33 = fma(be, V03, fma(de, V13, V33));
    R Qx0, Qx1, Qx2, Qx3, Qu0, Qu1, Quu00, Quu01, Quu10, Quu11;
    R M00, M01, M02, M03, M11, M12, M13, M22, M23, M33;
    Qx0 = t[10] + Vx0;
    Qx1 = t[11] + Vx1;
    if (kQR) {
      Qx2 = t[kT - 2] + fma(al, Vx0, fma(ga, Vx1, Vx2));
      Qx3 = t[kT - 1] + fma(be, Vx0, fma(de, Vx1, Vx3));
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
    } else {
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
    // ---- regularised inverse (iterative_ilqr.py:118-123): closed form first, the eigen-clamp only when needed
    R i00, i01, i11;
    {
      const R a = Quu00, b = (Quu01 + Quu10) / R(2.0), d = Quu11;
      const R aq = a + lamb, dq = d + lamb;
      const R r = rcp_fast(aq * dq - b * b);
      i00 = dq * r;
      i01 = -b * r;
      i11 = aq * r;
      if (!(a * d - b * b >= R(0.0) && a + d >= R(0.0))) quu_inverse(a, b, d, lamb, i00, i01, i11);
    }
    // The gains are kept with the opposite sign (kp = -k, Kp = -K: no negation instructions); the value-function
    // update below is invariant under that sign flip and the forward pass subtracts -- bit-identical results.
    const R k0 = fma(i01, Qu1, i00 * Qu0), k1 = fma(i11, Qu1, i01 * Qu0);
    const R K00 = fma(i01, H0, i00 * G0), K01 = fma(i01, H1, i00 * G1);
    const R K02 = fma(i01, X12, i00 * X02), K03 = fma(i01, X13, i00 * X03);
    const R K10 = fma(i11, H0, i01 * G0), K11 = fma(i11, H1, i01 * G1);
    const R K12 = fma(i11, X12, i01 * X02), K13 = fma(i11, X13, i01 * X03);
    {
      R* g = gains + (size_t)(10 * k) * S;
      g[0] = k0;
      g[S] = k1;
      g[2 * S] = K00;
      g[3 * S] = K01;
      g[4 * S] = K02;
      g[5 * S] = K03;
      g[6 * S] = K10;
      g[7 * S] = K11;
      g[8 * S] = K12;
      g[9 * S] = K13;
    }
    // ---- value function update with the unregularised Quu (iterative_ilqr.py:128-129)
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

// Full-step forward pass (iterative_ilqr.py:133-160), skewed: the trip for stage k evaluates sin/cos of heading k
// next to the feedback law of stage k.  Writes the trial trajectory to `dst`, returns cost_new (Jtarget as in
// ilqr_core.cuh forward_pass).
template <bool kQR, int S, typename R>
__device__ __forceinline__ R chain_forward(const KParamsT<R>& p, const CandT<R>& c, const R* cur, const R* gains, R* dst, R& Jtarget) {
  const int N = p.N;
  R x[4] = {c.x0[0], c.x0[1], c.x0[2], c.x0[3]};
  R Jn = R(0.0), Jt = R(0.0);
  R g[10], xo[4], uo[2];
#pragma unroll
  for (int i = 0; i < 10; ++i) g[i] = gains[i * S];
#pragma unroll
  for (int i = 0; i < 4; ++i) xo[i] = c.x0[i];
  uo[0] = cur[0];
  uo[1] = cur[S];
  for (int k = 0; k < N; ++k) {
    // operands of the next stage (clamped, consumed one trip later)
    R gn[10], xon[4], uon[2];
    {
      const int kn = k + 1 < N ? k + 1 : k;
      const R* gp = gains + (size_t)(10 * kn) * S;
#pragma unroll
      for (int i = 0; i < 10; ++i) gn[i] = gp[i * S];
      const R* cp = cur + (size_t)(8 * k) * S;
#pragma unroll
      for (int i = 0; i < 4; ++i) xon[i] = cp[(2 + i) * S];
      const R* cn = cur + (size_t)(8 * kn) * S;
      uon[0] = cn[0];
      uon[1] = cn[S];
    }
    // sin / cos of heading k: only the position update below needs them (stage 0 re-derives c.c0 / c.s0 rather
    // than branching, a branch would end the basic block the two chains share)
    R sn, cs;
    sincos_inline(x[3], &sn, &cs);
    const R d0 = x[0] - xo[0], d1 = x[1] - xo[1], d2 = x[2] - xo[2], d3 = x[3] - xo[3];
    // u + k + K (x - xo) with the sign-flipped gains of chain_riccati
    R ua = (uo[0] - g[0]) - fma(g[5], d3, fma(g[4], d2, fma(g[3], d1, g[2] * d0)));
    R ud = (uo[1] - g[1]) - fma(g[9], d3, fma(g[8], d2, fma(g[7], d1, g[6] * d0)));
    ua = clipd(ua, p.a_min, p.a_max);
    ud = clipd(ud, p.d_min, p.d_lim);
    if (kQR) {
      const R l_ctrl = quad2(p.Rm, ua, ud);
      Jn = Jn + quad4(p.Q, x[0] - c.xt[0], x[1] - c.xt[1], x[2] - c.xt[2], x[3] - c.xt[3]) + l_ctrl;
      Jt = Jt + quad4(p.Q, x[0] - p.xtarget[0], x[1] - p.xtarget[1], x[2] - p.xtarget[2], x[3] - p.xtarget[3]) + l_ctrl;
    }
    sincos_fixup(x[3], &sn, &cs);
    if (k > 0) {
      R* dp = dst + (size_t)(8 * (k - 1)) * S;
      dp[6 * S] = cs;
      dp[7 * S] = sn;
    }
    const R sd = x[2] * p.dt + ua * p.hdt2;
    x[0] = x[0] + cs * sd;
    x[1] = x[1] + sn * sd;
    x[2] = x[2] + ua * p.dt;
    x[3] = x[3] + ud * p.dt;
    {
      R* d = dst + (size_t)(8 * k) * S;
      d[0] = ua;
      d[S] = ud;
      d[2 * S] = x[0];
      d[3 * S] = x[1];
      d[4 * S] = x[2];
      d[5 * S] = x[3];
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) g[i] = gn[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) xo[i] = xon[i];
    uo[0] = uon[0];
    uo[1] = uon[1];
  }
  {
    R sn, cs;
    sincos_inline(x[3], &sn, &cs);
    sincos_fixup(x[3], &sn, &cs);
    R* dp = dst + (size_t)(8 * (N - 1)) * S;
    dp[6 * S] = cs;
    dp[7 * S] = sn;
  }
  const R T = terminal_cost(p, c, x);
  Jtarget = Jt + T;
  return Jn + T;
}

#ifdef __CUDACC__
// One candidate solve by a group of kGroup adjacent lanes (the first kRungs of them are the speculative lambda
// rungs, all of them share the stage tasks), as a small state machine that a warp drives convergently: every lane
// of the warp calls trip() each round (it contains warp-wide barriers and votes); groups without a live candidate
// pass through idle.  One trip consumes 1..kRungs reference iterations (iterative_ilqr.py:29-84), accepted in the
// reference's order.  Used by the controller-step kernels (kernels.cuh: chain_lat_kernel; spec_step.cuh).
template <typename R, bool kQR, int kRungs, int kGroup, int S>
struct GroupSolver {
  static constexpr unsigned kFull = 0xffffffffu;
  ChainLane<R, S> v;
  int rung, gbase_lane;
  bool is_rung;
  CandT<R> c;          // x0 / c0 / s0 / obstacle / xt / lamb are set by the caller before start()
  int wb;              // trial buffer this lane writes next
  const R* cur;        // current trajectory of the group: some lane's trial buffer
  bool dirty;          // the shared terms do not belong to `cur` yet
  bool running;
#ifdef I2LQR_TRIP_PROFILE
  long long prof[5] = {0, 0, 0, 0, 0};  // developer build: cycles in terms / riccati / forward / accept, rounds
#define I2_PROF(i)                  \
  {                                 \
    const long long t_ = clock64(); \
    prof[i] += t_ - tp_;            \
    tp_ = t_;                       \
  }
#else
#define I2_PROF(i)
#endif

  __device__ __forceinline__ void bind(R* smem, int lane, int N) {
    rung = lane % kGroup;
    is_rung = rung < kRungs;
    gbase_lane = lane - rung;
    v.col = smem + lane;
    v.grp = smem + gbase_lane;
    v.N = N;
    v.rungs = kGroup;
    wb = 0;
    cur = v.trial_of(0, 0);
    dirty = true;
    running = false;
    c = CandT<R>{};
  }
  // candidate start (utils/base.py:405-410 + the first rollout): zero inputs
  __device__ __forceinline__ void start(const KParamsT<R>& p) {
    c.J = chain_init<kQR, S>(p, c, rung == 0 ? v.trial(0) : nullptr);
    c.iters = 0;
    cur = v.trial_of(0, 0);
    wb = rung == 0 ? 1 : 0;
    dirty = true;
    running = true;
  }
  // One round for the whole warp.  Returns true on the lanes of a group whose candidate ended in this round; its
  // result is then `cur` (trajectory), c.lamb, c.iters, c.J.
  __device__ __forceinline__ bool trip(const KParamsT<R>& p) {
#ifdef I2LQR_TRIP_PROFILE
    long long tp_ = clock64();
    prof[4] += 1;
#endif
    __syncwarp();
    const bool run = running;
    if (run && dirty) chain_terms<kQR, S>(p, c, v, cur, rung);
    __syncwarp();
    I2_PROF(0)
    R Jn = R(0.0), Jt = R(0.0);
    if (run && is_rung) {
      R lamb_r = c.lamb;
      for (int r = 0; r < rung; ++r) lamb_r *= p.lamb_factor;  // the value r rejects lead to, same operations
      chain_riccati<kQR, S>(p, v, lamb_r);
      __syncwarp(__activemask());
      I2_PROF(1)
      Jn = chain_forward<kQR, S>(p, c, cur, v.gains(), v.trial(wb), Jt);
    }
    __syncwarp();
    I2_PROF(2)
    const unsigned acc_all = __ballot_sync(kFull, run && is_rung && Jn < c.J);
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
    const int src = winner >= 0 ? gbase_lane + winner : (int)(threadIdx.x & 31);
    const R Jn_w = __shfl_sync(kFull, Jn, src), Jt_w = __shfl_sync(kFull, Jt, src);
    const int wb_w = __shfl_sync(kFull, wb, src);
    if (run) {
      dirty = winner >= 0;
      if (winner >= 0) {
        cur = v.trial_of(winner, wb_w);  // the winner's trial is the current trajectory from here on
        if (rung == winner) wb ^= 1;
        if (rel_change_below(Jn_w, c.J, p.eps)) fin = true;
        c.J = kQR ? Jt_w : Jn_w;
      }
      c.lamb = lamb;
      c.iters += consumed;
      if (c.iters >= p.max_iter) fin = true;
      if (fin) running = false;
    }
    I2_PROF(3)
    return run && fin;
  }
  // Result in the reference's layout, xvar (4, N+1) and uvar (2, N), written by the lanes of the group together.
  __device__ __forceinline__ void store(double* xv, double* uv) const {
    const int N = v.N;
    for (int k = rung; k < N; k += kGroup) {
      const R* s = cur + (size_t)(8 * k) * S;
      uv[k] = (double)s[0];
      uv[N + k] = (double)s[S];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i * (N + 1) + k + 1] = (double)s[(2 + i) * S];
    }
    if (rung == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i * (N + 1)] = (double)c.x0[i];
    }
  }
  __device__ __forceinline__ void terminal_state(R (&xN)[4]) const {
    const R* s = cur + (size_t)(8 * (v.N - 1)) * S;
#pragma unroll
    for (int i = 0; i < 4; ++i) xN[i] = s[(2 + i) * S];
  }
};
#endif  // __CUDACC__

}  // namespace i2lqr
