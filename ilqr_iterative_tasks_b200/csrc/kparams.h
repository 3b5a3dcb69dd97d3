// kparams.h -- host-side derivation of the kernel parameter block (KParams) from the ABI's i2lqr_params.
#pragma once
#include <cmath>
#include <cstring>

#include "../../include/i2lqr.h"
#include "ilqr_core.cuh"

namespace i2lqr {

inline bool symmetric(const double* m, int n) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j)
      if (m[i * n + j] != m[j * n + i]) return false;
  return true;
}

inline void upper4(const double* m, double s, double* out) {
  int t = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = i; j < 4; ++j) out[t++] = s * m[i * 4 + j];
}


// Returns nullptr on success, else a static error message.
inline const char* fill_kparams(const i2lqr_params& p, KParams* k) {
  if (p.num_horizon < 2) return "num_horizon must be >= 2 (the reference's N==1 branch is unreachable, utils/base.py:438-450)";
  if (p.num_ss_points < 1 || p.num_ss_iter < 1) return "num_ss_points and num_ss_iter must be >= 1";
  if (!symmetric(p.matrix_Q, 4) || !symmetric(p.matrix_R, 2) || !symmetric(p.matrix_Qterminal, 4))
    return "matrix_Q, matrix_R and matrix_Qterminal must be symmetric";
  memset(k, 0, sizeof(*k));
  k->N = p.num_horizon;
  k->N_param = p.obstacle_horizon > 0 ? p.obstacle_horizon : p.num_horizon;
  k->K = p.num_ss_points;
  k->L = p.num_ss_iter;
  k->max_iter = p.max_ilqr_iter;
  k->max_relax = p.max_relax_iter;
  k->dt = p.timestep;
  k->dt2 = p.timestep * p.timestep;
  k->hdt2 = k->dt2 / 2.0;
  k->a_max = p.a_max;
  k->d_lim = p.delta_lim;
  k->a_min = -p.a_max;
  k->d_min = -p.delta_lim;
  k->eps = p.eps;
  k->lamb0 = p.lamb;
  k->lamb_factor = p.lamb_factor;
  k->max_lamb = p.max_lamb;
  k->cq2 = p.tuning_ctrl_q2;
  k->cg = p.tuning_ctrl_q1 * p.tuning_ctrl_q2;
  k->ch = p.tuning_ctrl_q1 * (p.tuning_ctrl_q2 * p.tuning_ctrl_q2);
  k->ca_lo = std::exp(-2.0 * p.tuning_ctrl_q2 * p.a_max);
  k->cd_lo = std::exp(-2.0 * p.tuning_ctrl_q2 * p.delta_lim);
  k->oq2 = p.tuning_obs_q2;
  k->og = p.tuning_obs_q1 * p.tuning_obs_q2;
  k->oh = p.tuning_obs_q1 * (p.tuning_obs_q2 * p.tuning_obs_q2);
  k->one_margin = 1 + p.safety_margin;
  for (int i = 0; i < 4; ++i) k->xtarget[i] = p.xtarget[i];
  upper4(p.matrix_Qterminal, 1.0, k->Qf);
  upper4(p.matrix_Qterminal, 2.0, k->Qf2);
  upper4(p.matrix_Q, 1.0, k->Q);
  upper4(p.matrix_Q, 2.0, k->Q2);
  k->Rm[0] = p.matrix_R[0];
  k->Rm[1] = p.matrix_R[1];
  k->Rm[2] = p.matrix_R[3];
  for (int i = 0; i < 3; ++i) k->R2[i] = 2.0 * k->Rm[i];
  bool qr = false;
  for (int i = 0; i < 16; ++i) qr = qr || p.matrix_Q[i] != 0.0;
  for (int i = 0; i < 4; ++i) qr = qr || p.matrix_R[i] != 0.0;
  k->has_qr = qr;
  return nullptr;
}

// Single-precision copy of the parameter block for the optional FP32 build.
inline KParamsT<float> to_float(const KParams& k) {
  KParamsT<float> f;
  f.N = k.N; f.N_param = k.N_param; f.K = k.K; f.L = k.L; f.max_iter = k.max_iter; f.max_relax = k.max_relax;
  f.has_qr = k.has_qr; f.pad0 = 0;
  f.dt = (float)k.dt; f.dt2 = (float)k.dt2; f.hdt2 = f.dt2 / 2.0f; f.a_max = (float)k.a_max; f.d_lim = (float)k.d_lim; f.a_min = (float)k.a_min; f.d_min = (float)k.d_min;
  f.eps = (float)k.eps; f.lamb0 = (float)k.lamb0; f.lamb_factor = (float)k.lamb_factor; f.max_lamb = (float)k.max_lamb;
  f.cg = (float)k.cg; f.ch = (float)k.ch; f.cq2 = (float)k.cq2; f.ca_lo = (float)k.ca_lo; f.cd_lo = (float)k.cd_lo;
  f.og = (float)k.og; f.oh = (float)k.oh; f.oq2 = (float)k.oq2; f.one_margin = (float)k.one_margin;
  for (int i = 0; i < 4; ++i) f.xtarget[i] = (float)k.xtarget[i];
  for (int i = 0; i < 10; ++i) { f.Qf[i] = (float)k.Qf[i]; f.Qf2[i] = (float)k.Qf2[i]; f.Q[i] = (float)k.Q[i]; f.Q2[i] = (float)k.Q2[i]; }
  for (int i = 0; i < 3; ++i) { f.Rm[i] = (float)k.Rm[i]; f.R2[i] = (float)k.R2[i]; }
  return f;
}

}  // namespace i2lqr
