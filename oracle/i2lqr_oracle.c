/*
 * i2lqr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar FP64 CPU restatement of the i2LQR per-timestep control solve of HybridRobotics/ilqr-iterative-tasks,
 * written from the reference's Python sources (citations are file:line relative to the reference root).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker or the timed CPU baseline.  The product library (libi2lqr.so) never links or calls it.
 *
 * Pin: tests/test_oracle.py checks every function here against the .npz files under tests/golden/, which are outputs of the
 * unmodified Python reference run in the build container (tests/golden/make_golden.py).
 *
 * Arithmetic: dense 4x4 / 4x2 / 2x2 products with sequential left-to-right sums, evaluated in the
 * association order of the reference's expressions; libm sin/cos/exp/sqrt.  The 2x2 symmetric eigen
 * decomposition behind the regularised inverse uses the Jacobi-rotation closed form in place
 * of LAPACK dgeev, the one place where the reference's arithmetic lives in an un-vendored dependency
 * (numpy.linalg.eig, iterative_ilqr/control/iterative_ilqr.py:118).
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/i2lqr.h"

#define NX 4
#define NU 2

/* systems/kinetic_bicycle.py:10-27 */
static void kinetic_bicycle(const double* x, const double* u, double dt, double* xn) {
  double s = x[2] * dt + (u[0] * (dt * dt)) / 2;
  xn[0] = x[0] + cos(x[3]) * s;
  xn[1] = x[1] + sin(x[3]) * s;
  xn[2] = x[2] + u[0] * dt;
  xn[3] = x[3] + u[1] * dt;
}

static double clip(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* v^T M v evaluated as (v^T M) @ v, n x n row-major M */
static double quad_form(const double* v, const double* M, int n) {
  double acc = 0;
  for (int j = 0; j < n; ++j) {
    double t = 0;
    for (int i = 0; i < n; ++i) t += v[i] * M[i * n + j];
    acc += t * v[j];
  }
  return acc;
}

/* control/ilqr_helper.py:59-64 applied to the obstacle ellipse (:32-51 stage, :118-149 terminal).
 * step = horizon index k for a stage, ilqr_param.num_horizon for the terminal call.  Adds into lx, lxx. */
static void obstacle_barrier_one(const i2lqr_params* p, const i2lqr_obstacle* o, const double* xk, int step, double* lx,
                                 double* lxx);

/* EXTENSION (obstacle lists, include/i2lqr.h i2lqr_set_obstacle_list): further obstacles of the calling thread's solve.  The
 * reference prices one Obstacle; a list is priced as the sum of its members' barriers, each evaluated by the reference's formula. */
static __thread const i2lqr_obstacle* g_extra = 0;
static __thread int g_n_extra = 0;
void orc_set_extra_obstacles(const i2lqr_obstacle* list, int n) {
  g_extra = list;
  g_n_extra = n;
}

static void obstacle_barrier(const i2lqr_params* p, const i2lqr_obstacle* o, const double* xk, int step, double* lx,
                             double* lxx) {
  obstacle_barrier_one(p, o, xk, step, lx, lxx);
  for (int m = 0; m < g_n_extra; ++m) obstacle_barrier_one(p, g_extra + m, xk, step, lx, lxx);
}

static void obstacle_barrier_one(const i2lqr_params* p, const i2lqr_obstacle* o, const double* xk, int step, double* lx,
                                 double* lxx) {
  if (!o || !o->present) return;
  double cx = o->x, cy = o->y;
  if (o->moving_option == 1) cy = o->y + step * o->spd;
  if (o->moving_option == 2) cx = o->x - step * o->spd;
  double d0 = xk[0] - cx, d1 = xk[1] - cy;
  double P0 = 1 / (o->width * o->width), P1 = 1 / (o->height * o->height);
  double h = 1 + p->safety_margin - ((d0 * P0) * d0 + (d1 * P1) * d1);
  double hd[NX] = {(-2 * P0) * d0, (-2 * P1) * d1, 0, 0};
  double q1 = p->tuning_obs_q1, q2 = p->tuning_obs_q2;
  double e = exp(q2 * h);
  double c1 = q1 * q2 * e, c2 = q1 * (q2 * q2) * e;
  for (int i = 0; i < NX; ++i) lx[i] += c1 * hd[i];
  for (int i = 0; i < NX; ++i)
    for (int j = 0; j < NX; ++j) lxx[i * NX + j] += c2 * (hd[i] * hd[j]);
}

/* control/ilqr_helper.py:83-103: four exponential input barriers */
static void control_barrier(const i2lqr_params* p, const double* u, double* bd, double* bdd) {
  double q1 = p->tuning_ctrl_q1, q2 = p->tuning_ctrl_q2;
  double c1 = u[0] - p->a_max, c2 = -p->a_max - u[0];
  double c3 = u[1] - p->delta_lim, c4 = -p->delta_lim - u[1];
  double e1 = exp(q2 * c1), e2 = exp(q2 * c2), e3 = exp(q2 * c3), e4 = exp(q2 * c4);
  double g = q1 * q2, hh = q1 * (q2 * q2);
  bd[0] = g * e1 * 1.0 + g * e2 * -1.0;
  bd[1] = g * e3 * 1.0 + g * e4 * -1.0;
  bdd[0] = hh * e1 + hh * e2;
  bdd[1] = 0;
  bdd[2] = 0;
  bdd[3] = hh * e3 + hh * e4;
}

/* Regularised inverse of the 2x2 Quu: eig, clamp negative eigenvalues to 0, add lamb, V diag(1/e) V^T
 * (control/iterative_ilqr.py:118-123); Jacobi-rotation closed form of the symmetric eigenproblem. */
static void quu_inverse(const double* Quu, double lamb, double* inv) {
  double a = Quu[0], d = Quu[3], b = (Quu[1] + Quu[2]) / 2;
  double c, s, e1, e2;
  if (b == 0.0) {
    c = 1;
    s = 0;
    e1 = a;
    e2 = d;
  } else {
    double tau = (d - a) / (2 * b);
    double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1 + tau * tau));
    c = 1 / sqrt(1 + t * t);
    s = t * c;
    e1 = a - t * b;
    e2 = d + t * b;
  }
  if (e1 < 0) e1 = 0.0;
  if (e2 < 0) e2 = 0.0;
  e1 += lamb;
  e2 += lamb;
  double r1 = 1.0 / e1, r2 = 1.0 / e2;
  /* eigenvectors (c,-s) and (s,c) */
  inv[0] = c * r1 * c + s * r2 * s;
  inv[1] = c * r1 * -s + s * r2 * c;
  inv[2] = inv[1];
  inv[3] = -s * r1 * -s + c * r2 * c;
}

static void matmul(const double* A, const double* B, double* C, int r, int c, int pcols) { /* (r x c)(c x p) */
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < pcols; ++j) {
      double t = 0;
      for (int k = 0; k < c; ++k) t += A[i * c + k] * B[k * pcols + j];
      C[i * pcols + j] = t;
    }
}
static void transpose(const double* A, double* At, int r, int c) {
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) At[j * r + i] = A[i * c + j];
}

/* control/iterative_ilqr.py:88-130.  X (N+1,4), U (N,2), dX (N+1,4) time-major; gains k (N,2), K (N,2,4). */
static void backward_pass(const i2lqr_params* p, const i2lqr_obstacle* o, int N, const double* X, const double* U,
                          const double* dX, const double* xterm, double lamb, double* gk, double* gK) {
  double dt = p->timestep;
  double Vx[NX], Vxx[NX * NX];
  /* get_cost_final, control/ilqr_helper.py:106-150 */
  {
    double diff[NX];
    for (int i = 0; i < NX; ++i) diff[i] = X[N * NX + i] - xterm[i];
    for (int i = 0; i < NX; ++i) {
      double t = 0;
      for (int j = 0; j < NX; ++j) t += (2 * p->matrix_Qterminal[i * NX + j]) * diff[j];
      Vx[i] = t;
    }
    for (int i = 0; i < NX * NX; ++i) Vxx[i] = 2 * p->matrix_Qterminal[i];
    obstacle_barrier(p, o, &X[N * NX], p->obstacle_horizon > 0 ? p->obstacle_horizon : p->num_horizon, Vx, Vxx);
  }
  for (int k = N - 1; k >= 0; --k) {
    /* A, B at (v_{k+1}, theta_{k+1}, a_k): control/iterative_ilqr.py:92-99, systems/kinetic_bicycle.py:30-52 */
    double v = X[(k + 1) * NX + 2], th = X[(k + 1) * NX + 3], acc = U[k * NU + 0];
    double sp = v * dt + (acc * (dt * dt)) / 2;
    double A[NX * NX] = {1, 0, cos(th) * dt, -sp * sin(th), 0, 1, sin(th) * dt, sp * cos(th), 0, 0, 1, 0, 0, 0, 0, 1};
    double B[NX * NU] = {(dt * dt) * cos(th) / 2, 0, (dt * dt) * sin(th) / 2, 0, dt, 0, 0, dt};
    /* get_cost_derivation, control/ilqr_helper.py:25-55 */
    double lu[NU], luu[NU * NU], lx[NX], lxx[NX * NX], bd[NU], bdd[NU * NU];
    control_barrier(p, &U[k * NU], bd, bdd);
    for (int i = 0; i < NU; ++i) {
      double t = 0;
      for (int j = 0; j < NU; ++j) t += (2 * p->matrix_R[i * NU + j]) * U[k * NU + j];
      lu[i] = t + bd[i];
    }
    for (int i = 0; i < NU * NU; ++i) luu[i] = 2 * p->matrix_R[i] + bdd[i];
    for (int i = 0; i < NX; ++i) {
      double t = 0;
      for (int j = 0; j < NX; ++j) t += (2 * p->matrix_Q[i * NX + j]) * dX[k * NX + j];
      lx[i] = t;
    }
    for (int i = 0; i < NX * NX; ++i) lxx[i] = 2 * p->matrix_Q[i];
    obstacle_barrier(p, o, &X[k * NX], k, lx, lxx);
    /* Riccati step, control/iterative_ilqr.py:112-129 */
    double At[NX * NX], Bt[NU * NX], Qx[NX], Qu[NU], Qxx[NX * NX], Quu[NU * NU], Qux[NU * NX];
    double AtV[NX * NX], BtV[NU * NX], tmp[NX * NX];
    transpose(A, At, NX, NX);
    transpose(B, Bt, NX, NU);
    matmul(At, Vx, Qx, NX, NX, 1);
    for (int i = 0; i < NX; ++i) Qx[i] = lx[i] + Qx[i];
    matmul(Bt, Vx, Qu, NU, NX, 1);
    for (int i = 0; i < NU; ++i) Qu[i] = lu[i] + Qu[i];
    matmul(At, Vxx, AtV, NX, NX, NX);
    matmul(AtV, A, tmp, NX, NX, NX);
    for (int i = 0; i < NX * NX; ++i) Qxx[i] = lxx[i] + tmp[i];
    matmul(Bt, Vxx, BtV, NU, NX, NX);
    matmul(BtV, B, tmp, NU, NX, NU);
    for (int i = 0; i < NU * NU; ++i) Quu[i] = luu[i] + tmp[i];
    matmul(BtV, A, Qux, NU, NX, NX);
    double inv[NU * NU];
    quu_inverse(Quu, lamb, inv);
    double ninv[NU * NU];
    for (int i = 0; i < NU * NU; ++i) ninv[i] = -inv[i];
    double* kk = &gk[k * NU];
    double* KK = &gK[k * NU * NX];
    matmul(ninv, Qu, kk, NU, NU, 1);
    matmul(ninv, Qux, KK, NU, NU, NX);
    double Kt[NX * NU], KtQ[NX * NU], t1[NX], t2[NX * NX];
    transpose(KK, Kt, NU, NX);
    matmul(Kt, Quu, KtQ, NX, NU, NU);
    matmul(KtQ, kk, t1, NX, NU, 1);
    matmul(KtQ, KK, t2, NX, NU, NX);
    for (int i = 0; i < NX; ++i) Vx[i] = Qx[i] - t1[i];
    for (int i = 0; i < NX * NX; ++i) Vxx[i] = Qxx[i] - t2[i];
  }
}

/* control/iterative_ilqr.py:133-160 */
static double forward_pass(const i2lqr_params* p, int N, const double* X, const double* U, const double* xterm,
                           const double* gk, const double* gK, double* Xn, double* Un) {
  double cost = 0;
  memcpy(Xn, X, NX * sizeof(double));
  for (int k = 0; k < N; ++k) {
    double dx[NX];
    for (int i = 0; i < NX; ++i) dx[i] = Xn[k * NX + i] - X[k * NX + i];
    for (int i = 0; i < NU; ++i) {
      double t = 0;
      for (int j = 0; j < NX; ++j) t += gK[(k * NU + i) * NX + j] * dx[j];
      Un[k * NU + i] = U[k * NU + i] + gk[k * NU + i] + t;
    }
    Un[k * NU + 0] = clip(Un[k * NU + 0], -p->a_max, p->a_max);
    Un[k * NU + 1] = clip(Un[k * NU + 1], -p->delta_lim, p->delta_lim);
    kinetic_bicycle(&Xn[k * NX], &Un[k * NU], p->timestep, &Xn[(k + 1) * NX]);
    double e[NX];
    for (int i = 0; i < NX; ++i) e[i] = Xn[k * NX + i] - xterm[i];
    double l_state = quad_form(e, p->matrix_Q, NX);
    double l_ctrl = quad_form(&Un[k * NU], p->matrix_R, NU);
    cost = cost + l_state + l_ctrl;
  }
  double e[NX];
  for (int i = 0; i < NX; ++i) e[i] = Xn[N * NX + i] - xterm[i];
  cost = cost + quad_form(e, p->matrix_Qterminal, NX);
  return cost;
}

/* control/iterative_ilqr.py:7-85 with the candidate initialisation of utils/base.py:405-410.
 * uvar_out (2,N), xvar_out (4,N+1) in the reference's layout (time is the last axis). */
int orc_ilqr_ex(const i2lqr_params* p, const i2lqr_obstacle* o, const double* x0, const double* xterm, double lamb,
                const double* uvar_init, double* uvar_out, double* xvar_out, double* lamb_out, int32_t* iters_out) {
  int N = p->num_horizon;
  size_t nd = (size_t)(N + 1) * NX * 3 + (size_t)N * NU * 3 + (size_t)N * NU * NX;
  double* mem = (double*)calloc(nd, sizeof(double));
  if (!mem) return 1;
  double *X = mem, *Xn = X + (N + 1) * NX, *dX = Xn + (N + 1) * NX, *U = dX + (N + 1) * NX, *Un = U + N * NU,
         *gk = Un + N * NU, *gK = gk + N * NU;
  for (int i = 0; i < NX; ++i) {
    X[i] = x0[i];
    dX[i] = x0[i] - p->xtarget[i];
  }
  if (uvar_init) /* warm start of a direct ilqr() call; calc_input passes zeros (utils/base.py:405) */
    for (int k = 0; k < N; ++k)
      for (int i = 0; i < NU; ++i) U[k * NU + i] = uvar_init[i * N + k];
  int iters = 0;
  for (int it = 0; it < p->max_ilqr_iter; ++it) {
    double cost = 0;
    for (int k = 0; k < N; ++k) {
      U[k * NU + 0] = clip(U[k * NU + 0], -p->a_max, p->a_max);
      U[k * NU + 1] = clip(U[k * NU + 1], -p->delta_lim, p->delta_lim);
      kinetic_bicycle(&X[k * NX], &U[k * NU], p->timestep, &X[(k + 1) * NX]);
      double e[NX];
      for (int i = 0; i < NX; ++i) {
        dX[(k + 1) * NX + i] = X[(k + 1) * NX + i] - p->xtarget[i];
        e[i] = X[k * NX + i] - p->xtarget[i];
      }
      double l_state = quad_form(e, p->matrix_Q, NX);
      double l_ctrl = quad_form(&U[k * NU], p->matrix_R, NU);
      cost = cost + l_state + l_ctrl;
    }
    double e[NX];
    for (int i = 0; i < NX; ++i) e[i] = X[N * NX + i] - xterm[i];
    cost = cost + quad_form(e, p->matrix_Qterminal, NX);
    backward_pass(p, o, N, X, U, dX, xterm, lamb, gk, gK);
    double cost_new = forward_pass(p, N, X, U, xterm, gk, gK, Xn, Un);
    ++iters;
    if (cost_new < cost) {
      memcpy(X, Xn, (size_t)(N + 1) * NX * sizeof(double));
      memcpy(U, Un, (size_t)N * NU * sizeof(double));
      lamb /= p->lamb_factor;
      if (fabs((cost_new - cost) / cost) < p->eps) break;
    } else {
      lamb *= p->lamb_factor;
      if (lamb > p->max_lamb) break;
    }
  }
  if (uvar_out)
    for (int k = 0; k < N; ++k)
      for (int i = 0; i < NU; ++i) uvar_out[i * N + k] = U[k * NU + i];
  if (xvar_out)
    for (int k = 0; k <= N; ++k)
      for (int i = 0; i < NX; ++i) xvar_out[i * (N + 1) + k] = X[k * NX + i];
  if (lamb_out) *lamb_out = lamb;
  if (iters_out) *iters_out = iters;
  free(mem);
  return 0;
}

int orc_ilqr(const i2lqr_params* p, const i2lqr_obstacle* o, const double* x0, const double* xterm, double lamb,
             double* uvar_out, double* xvar_out, double* lamb_out, int32_t* iters_out) {
  return orc_ilqr_ex(p, o, x0, xterm, lamb, NULL, uvar_out, xvar_out, lamb_out, iters_out);
}

/* B independent problems on `nthreads` host threads (pthreads, dynamic chunks of 64): the CPU baseline of
 * bench.py.  (This image's gcc has no libgomp, hence no OpenMP.) */
typedef struct {
  const i2lqr_params* p;
  int64_t B;
  const double *x0, *x_term, *lamb_in, *u_init;
  const i2lqr_obstacle* obs;
  int64_t n_obs;
  double *uvar, *xvar, *lamb_out;
  int32_t* iters;
  atomic_llong next;
  atomic_int bad;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  int N = j->p->num_horizon;
  for (;;) {
    int64_t b0 = atomic_fetch_add(&j->next, 64);
    if (b0 >= j->B) break;
    int64_t b1 = b0 + 64 < j->B ? b0 + 64 : j->B;
    for (int64_t b = b0; b < b1; ++b) {
      const i2lqr_obstacle* o = j->n_obs == 0 ? NULL : (j->n_obs == 1 ? j->obs : j->obs + b);
      int r = orc_ilqr_ex(j->p, o, j->x0 + b * NX, j->x_term + b * NX, j->lamb_in ? j->lamb_in[b] : j->p->lamb,
                       j->u_init ? j->u_init + b * NU * N : NULL, j->uvar ? j->uvar + b * NU * N : NULL, j->xvar ? j->xvar + b * NX * (N + 1) : NULL,
                       j->lamb_out ? j->lamb_out + b : NULL, j->iters ? j->iters + b : NULL);
      if (r) atomic_store(&j->bad, 1);
    }
  }
  return NULL;
}

int orc_ilqr_batch(const i2lqr_params* p, int64_t B, const double* x0, const double* x_term, const double* lamb_in,
                   const i2lqr_obstacle* obs, int64_t n_obs, const double* u_init, double* uvar, double* xvar,
                   double* lamb_out, int32_t* iters, int nthreads) {
  batch_job job = {p, B, x0, x_term, lamb_in, u_init, obs, n_obs, uvar, xvar, lamb_out, iters, 0, 0};
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, batch_worker, &job);
  batch_worker(&job);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
  return atomic_load(&job.bad);
}

/* numpy.argsort(d) with the default kind, as the reference calls it (utils/base.py:339).  The arithmetic lives in an
 * un-vendored dependency (requirements.txt pins numpy==1.19.5): its scalar `aquicksort` is an introsort -- median-of-3
 * partition while a range holds more than 16 elements, insertion sort below, heapsort once the depth budget
 * 2*floor(log2 n) is spent.  The order it leaves EQUAL keys in is a deterministic function of the input and matters
 * here (the initial lap ends in duplicated rows).  Restated from the published algorithm and checked
 * against NumPy 2.3.5 executed with its SIMD dispatch switched off (tests/golden/make_golden.py: knn.npz, every
 * steps_*.npz); NumPy >= 1.25 on AVX2 / AVX-512 hosts dispatches float64 argsort to x86-simd-sort, whose tie order is
 * hardware-specific and is not what the pinned version does. */
static int dless(double a, double b) { return a < b || (b != b && a == a); }
static void np_aheapsort(const double* v, int64_t* tosort, int64_t n) {
  int64_t* a = tosort - 1; /* 1-based */
  int64_t i, j, l, tmp;
  for (l = n >> 1; l > 0; --l) {
    tmp = a[l];
    for (i = l, j = l << 1; j <= n;) {
      if (j < n && dless(v[a[j]], v[a[j + 1]])) j += 1;
      if (dless(v[tmp], v[a[j]])) {
        a[i] = a[j];
        i = j;
        j += j;
      } else
        break;
    }
    a[i] = tmp;
  }
  for (; n > 1;) {
    tmp = a[n];
    a[n] = a[1];
    n -= 1;
    for (i = 1, j = 2; j <= n;) {
      if (j < n && dless(v[a[j]], v[a[j + 1]])) j++;
      if (dless(v[tmp], v[a[j]])) {
        a[i] = a[j];
        i = j;
        j += j;
      } else
        break;
    }
    a[i] = tmp;
  }
}
#define NP_SMALL_QUICKSORT 15 /* ranges of up to 16 elements go to the insertion sort */
void orc_np_argsort(const double* v, int64_t num, int64_t* tosort) {
  for (int64_t j = 0; j < num; ++j) tosort[j] = j;
  if (num < 2) return;
  int64_t *pl = tosort, *pr = tosort + num - 1, *stack[128], **sptr = stack, *pm, *pi, *pj, *pk, vi, tmp;
  int depth[128], *psdepth = depth, cdepth = 0;
  for (int64_t n = num; n > 1; n >>= 1) cdepth += 2; /* 2 * msb(num) */
  double vp;
#define ISWAP(a, b) (tmp = (a), (a) = (b), (b) = tmp)
  for (;;) {
    if (cdepth < 0) {
      np_aheapsort(v, pl, pr - pl + 1);
    } else {
      while (pr - pl > NP_SMALL_QUICKSORT) {
        pm = pl + ((pr - pl) >> 1);
        if (dless(v[*pm], v[*pl])) ISWAP(*pm, *pl);
        if (dless(v[*pr], v[*pm])) ISWAP(*pr, *pm);
        if (dless(v[*pm], v[*pl])) ISWAP(*pm, *pl);
        vp = v[*pm];
        pi = pl;
        pj = pr - 1;
        ISWAP(*pm, *pj);
        for (;;) {
          do ++pi;
          while (dless(v[*pi], vp));
          do --pj;
          while (dless(vp, v[*pj]));
          if (pi >= pj) break;
          ISWAP(*pi, *pj);
        }
        pk = pr - 1;
        ISWAP(*pi, *pk);
        if (pi - pl < pr - pi) { /* larger part on the stack */
          *sptr++ = pi + 1;
          *sptr++ = pr;
          pr = pi - 1;
        } else {
          *sptr++ = pl;
          *sptr++ = pi - 1;
          pl = pi + 1;
        }
        *psdepth++ = --cdepth;
      }
      for (pi = pl + 1; pi <= pr; ++pi) {
        vi = *pi;
        vp = v[vi];
        pj = pi;
        pk = pi - 1;
        while (pj > pl && dless(vp, v[*pk])) *pj-- = *pk--;
        *pj = vi;
      }
    }
    if (sptr == stack) break;
    pr = *(--sptr);
    pl = *(--sptr);
    cdepth = *(--psdepth);
  }
#undef ISWAP
}

/* iLqr.select_close_ss (utils/base.py:332-341): 1-norm distances, numpy.argsort, first K. */
int orc_select_knn(const double* lap, int T, const double* x0, int K, int64_t* idx_out) {
  if (T < 1) return 0;
  double* d = (double*)calloc((size_t)T, sizeof(double));
  int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)T);
  for (int j = 0; j < T; ++j)
    d[j] = ((fabs(lap[j * NX + 0] - x0[0]) + fabs(lap[j * NX + 1] - x0[1])) + fabs(lap[j * NX + 2] - x0[2])) +
           fabs(lap[j * NX + 3] - x0[3]);
  orc_np_argsort(d, T, order);
  int n = K < T ? K : T;
  for (int j = 0; j < n; ++j) idx_out[j] = order[j];
  free(d);
  free(order);
  return n;
}

/* utils/base.py:427-437: relax-cost ladder; `it` is the outer iteration (threshold 80 i / 10^it). */
double orc_relax_cost(const i2lqr_params* p, const double* xN, const double* xterm, double qfun, int it) {
  double s = 0;
  for (int i = 0; i < NX; ++i) s += (xN[i] - xterm[i]) * (xN[i] - xterm[i]);
  double r = sqrt(s);
  double scale = 1;
  for (int i = 0; i < it; ++i) scale *= 10;
  for (int i = 1; i <= p->max_relax_iter; ++i) {
    if (r <= 80.0 * i / scale) return qfun + p->num_horizon + 100 * i;
    if (r > 80.0 * p->max_relax_iter / scale) return INFINITY;
  }
  return NAN; /* only reachable when r is NaN (the reference would reuse a stale cost) */
}

/* Python list comparison a < b on cost lists (utils/base.py:462) */
static int list_less(const double* a, int na, const double* b, int nb) {
  int n = na < nb ? na : nb;
  for (int i = 0; i < n; ++i)
    if (a[i] != b[i]) return a[i] < b[i];
  return na < nb;
}

/* Solve branch of iLqr.calc_input (utils/base.py:384-478) for one controller.
 * laps: n_laps pointers to (T,4) arrays (all stored laps, oldest first), lap_len their lengths.
 * Outputs: u (2), x_pred (4,N+1), u_pred (2,N), winner {lap (absolute), rank, idx, cost}, shrink flag.
 * Optional trace arrays shaped (3, L, K): sel_idx (-1 unused), cand_cost, cand_lamb, cand_iters. */
int orc_calc_input(const i2lqr_params* p, const i2lqr_obstacle* o, int n_laps, const double* const* laps,
                   const int32_t* lap_len, const double* x, double* u, double* x_pred, double* u_pred,
                   int32_t* winner_lap, int32_t* winner_rank, int64_t* winner_idx, double* winner_cost,
                   int32_t* shrink, int64_t* sel_idx, double* cand_cost, double* cand_lamb, int32_t* cand_iters) {
  int N = p->num_horizon, K = p->num_ss_points, L = p->num_ss_iter;
  int min_iter = n_laps - L > 0 ? n_laps - L : 0;
  int n_act = n_laps - min_iter;
  size_t xsz = (size_t)NX * (N + 1), usz = (size_t)NU * N;
  double* xs = (double*)malloc(sizeof(double) * xsz * (size_t)n_act * K);
  double* us = (double*)malloc(sizeof(double) * usz * (size_t)n_act * K);
  double* costs = (double*)malloc(sizeof(double) * (size_t)n_act * K);
  int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_act * K);
  int* cnt = (int*)malloc(sizeof(int) * (size_t)n_act);
  double* xp = (double*)calloc(xsz, sizeof(double));
  double* up = (double*)calloc(usz, sizeof(double));
  int best_l = 0, best_r = 0;
  if (sel_idx)
    for (int i = 0; i < I2LQR_OUTER_ITERS * L * K; ++i) {
      sel_idx[i] = -1;
      if (cand_cost) cand_cost[i] = NAN;
      if (cand_lamb) cand_lamb[i] = NAN;
      if (cand_iters) cand_iters[i] = -1;
    }
  for (int it = 0; it < I2LQR_OUTER_ITERS; ++it) {
    double xg[NX];
    for (int i = 0; i < NX; ++i) xg[i] = it == 0 ? x[i] : xp[i * (N + 1) + N];
    for (int l = 0; l < n_act; ++l) {
      const double* lap = laps[min_iter + l];
      int T = lap_len[min_iter + l];
      double lamb = p->lamb;
      cnt[l] = orc_select_knn(lap, T, xg, K, idx + (size_t)l * K);
      for (int r = 0; r < cnt[l]; ++r) {
        int64_t j = idx[(size_t)l * K + r];
        double* xv = xs + ((size_t)l * K + r) * xsz;
        double* uv = us + ((size_t)l * K + r) * usz;
        int32_t iters = 0;
        orc_ilqr(p, o, x, lap + j * NX, lamb, uv, xv, &lamb, &iters);
        double xN[NX];
        for (int i = 0; i < NX; ++i) xN[i] = xv[i * (N + 1) + N];
        double c = orc_relax_cost(p, xN, lap + j * NX, (double)(T - 1 - j), it);
        costs[(size_t)l * K + r] = c;
        size_t t = ((size_t)it * L + l) * K + r;
        if (sel_idx) sel_idx[t] = j;
        if (cand_cost) cand_cost[t] = c;
        if (cand_lamb) cand_lamb[t] = lamb;
        if (cand_iters) cand_iters[t] = iters;
      }
    }
    best_l = 0;
    for (int l = 1; l < n_act; ++l)
      if (list_less(costs + (size_t)l * K, cnt[l], costs + (size_t)best_l * K, cnt[best_l])) best_l = l;
    best_r = 0;
    for (int r = 1; r < cnt[best_l]; ++r)
      if (costs[(size_t)best_l * K + r] < costs[(size_t)best_l * K + best_r]) best_r = r;
    memcpy(xp, xs + ((size_t)best_l * K + best_r) * xsz, sizeof(double) * xsz);
    memcpy(up, us + ((size_t)best_l * K + best_r) * usz, sizeof(double) * usz);
  }
  int64_t widx = idx[(size_t)best_l * K + best_r];
  if (u) {
    u[0] = up[0];
    u[1] = up[N];
  }
  if (x_pred) memcpy(x_pred, xp, sizeof(double) * xsz);
  if (u_pred) memcpy(u_pred, up, sizeof(double) * usz);
  if (winner_lap) *winner_lap = min_iter + best_l;
  if (winner_rank) *winner_rank = best_r;
  if (winner_idx) *winner_idx = widx;
  if (winner_cost) *winner_cost = costs[(size_t)best_l * K + best_r];
  if (shrink) *shrink = (widx + 1) > (lap_len[min_iter + best_l] - 1);
  free(xs);
  free(us);
  free(costs);
  free(idx);
  free(cnt);
  free(xp);
  free(up);
  return 0;
}

/* backward_pass / forward_pass in the reference's own array layout (time on the last axis): xvar (4,N+1),
 * uvar (2,N), dX (4,N+1), k (2,N), K (2,4,N).  control/iterative_ilqr.py:88-160. */
static void to_time_major(const double* a, int rows, int cols, double* out) { /* (rows, cols) -> (cols, rows) */
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) out[c * rows + r] = a[r * cols + c];
}
int orc_backward_pass(const i2lqr_params* p, const i2lqr_obstacle* o, const double* xvar, const double* uvar,
                      const double* dX, const double* xterm, double lamb, double* k_out, double* K_out) {
  int N = p->num_horizon;
  double* mem = (double*)malloc(sizeof(double) * (size_t)(2 * NX * (N + 1) + NU * N + NU * N + NU * NX * N));
  if (!mem) return 1;
  double *X = mem, *D = X + NX * (N + 1), *U = D + NX * (N + 1), *gk = U + NU * N, *gK = gk + NU * N;
  to_time_major(xvar, NX, N + 1, X);
  to_time_major(dX, NX, N + 1, D);
  to_time_major(uvar, NU, N, U);
  backward_pass(p, o, N, X, U, D, xterm, lamb, gk, gK);
  for (int k = 0; k < N; ++k) {
    for (int i = 0; i < NU; ++i) k_out[i * N + k] = gk[k * NU + i];
    for (int i = 0; i < NU * NX; ++i) K_out[i * N + k] = gK[k * NU * NX + i];
  }
  free(mem);
  return 0;
}
int orc_forward_pass(const i2lqr_params* p, const double* xvar, const double* uvar, const double* xterm,
                     const double* k_in, const double* K_in, double* xvar_new, double* uvar_new, double* cost_new) {
  int N = p->num_horizon;
  double* mem = (double*)malloc(sizeof(double) * (size_t)(2 * NX * (N + 1) + 2 * NU * N + NU * N + NU * NX * N));
  if (!mem) return 1;
  double *X = mem, *Xn = X + NX * (N + 1), *U = Xn + NX * (N + 1), *Un = U + NU * N, *gk = Un + NU * N, *gK = gk + NU * N;
  to_time_major(xvar, NX, N + 1, X);
  to_time_major(uvar, NU, N, U);
  for (int k = 0; k < N; ++k) {
    for (int i = 0; i < NU; ++i) gk[k * NU + i] = k_in[i * N + k];
    for (int i = 0; i < NU * NX; ++i) gK[k * NU * NX + i] = K_in[i * N + k];
  }
  *cost_new = forward_pass(p, N, X, U, xterm, gk, gK, Xn, Un);
  for (int k = 0; k <= N; ++k)
    for (int i = 0; i < NX; ++i) xvar_new[i * (N + 1) + k] = Xn[k * NX + i];
  for (int k = 0; k < N; ++k)
    for (int i = 0; i < NU; ++i) uvar_new[i * N + k] = Un[k * NU + i];
  free(mem);
  return 0;
}

/* systems/kinetic_bicycle.py:10-27, exported for the plant-step checks */
void orc_kinetic_bicycle(const double* x, const double* u, double dt, double* xn) { kinetic_bicycle(x, u, dt, xn); }

void orc_default_params(i2lqr_params* p) {
  memset(p, 0, sizeof(*p));
  p->num_horizon = 6;
  p->num_ss_points = 8;
  p->num_ss_iter = 1;
  p->max_ilqr_iter = 150;
  p->max_relax_iter = 55;
  p->max_sim_steps = 121;
  p->matrix_Qterminal[0] = 2 * 1.0;
  p->matrix_Qterminal[5] = 2 * 1.0;
  p->matrix_Qterminal[10] = 2 * 20.0;
  p->matrix_Qterminal[15] = 2 * 0.02;
  p->tuning_ctrl_q1 = 1.0;
  p->tuning_ctrl_q2 = 1.0;
  p->tuning_obs_q1 = 2.74;
  p->tuning_obs_q2 = 2.74;
  p->safety_margin = 0.0;
  p->eps = 1e-2;
  p->lamb = 1;
  p->lamb_factor = 10;
  p->max_lamb = 1000;
  p->a_max = 2.0;
  p->delta_lim = 1.57;
  p->timestep = 1.0;
  p->goal_tol = 0.8;
}
