/*
 * i2lqr.h -- C ABI of the B200-native i2LQR per-timestep control solve.
 *
 * The reference (HybridRobotics/ilqr-iterative-tasks) has no FFI: its boundary for this path is two plain
 * Python call levels.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference root).  Signatures carry plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; i2lqr_last_error() gives the message of
 *     the last failure on the calling thread.  The Python shim turns a failure into RuntimeError, the one
 *     exception the reference's caller handles (iterative_ilqr/utils/base.py:146-155).
 *   - all floating point is IEEE double; matrices are row-major; arrays use the reference's shapes:
 *     xvar (4, N+1), uvar (2, N) per candidate, laps (T, 4) as passed to add_trajectory.
 *   - "host" pointers are ordinary host memory (pinned or not); "_device" entry points take device
 *     pointers resident on the handle's GPU and a cudaStream_t passed as void*.
 *   - there is no CPU fallback: every compute entry point fails if no sm_100-class GPU is usable.
 */
#ifndef I2LQR_H_
#define I2LQR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I2LQR_ABI_VERSION 1
#define I2LQR_X_DIM 4 /* iterative_ilqr/utils/constants_kinetic_bicycle.py:2 */
#define I2LQR_U_DIM 2
#define I2LQR_OUTER_ITERS 3 /* calc_input breaks at iter == 2: iterative_ilqr/utils/base.py:472-478 */

/* iLqrParam (iterative_ilqr/utils/base.py:242-302) + KineticBicycleParam (:15-20) + the constants the
 * simulator hard-codes (Simulator.sim :708-726), flattened into one POD. */
typedef struct i2lqr_params {
  int32_t num_horizon;    /* N, iLqrParam.num_horizon (6) */
  int32_t num_ss_points;  /* K, iLqrParam.num_ss_points (8) */
  int32_t num_ss_iter;    /* L, iLqrParam.num_ss_iter (1) */
  int32_t max_ilqr_iter;  /* 150 */
  int32_t max_relax_iter; /* 55 */
  int32_t max_sim_steps;  /* 121, Simulator.sim :709-710 */
  int32_t obstacle_horizon; /* steps the obstacle is advanced for the terminal barrier: the reference uses
                               ilqr_param.num_horizon there (control/ilqr_helper.py:134-139) even when ilqr() is
                               called with another num_horizon; 0 = same as num_horizon */
  int32_t precision;      /* 0 = FP64 (the product path, parity <= 1e-9); 1 = optional FP32 build (BASELINE config 4:
                             <= 1e-4 against FP64 on single solves whose discrete path matches); I/O stays double */
  double matrix_Q[16];         /* default 0 */
  double matrix_R[4];          /* default 0 */
  double matrix_Qterminal[16]; /* default 2*diag(1,1,20,0.02) */
  double tuning_ctrl_q1, tuning_ctrl_q2; /* input barriers, 1, 1 */
  double tuning_obs_q1, tuning_obs_q2;   /* obstacle barrier, 2.74, 2.74 */
  double safety_margin;                  /* 0 */
  double eps;                            /* 1e-2 */
  double lamb, lamb_factor, max_lamb;    /* 1, 10, 1000 */
  double a_max;     /* KineticBicycleParam.a_max (2.0) */
  double delta_lim; /* round(KineticBicycleParam.delta_max, 2) = 1.57, iterative_ilqr/control/iterative_ilqr.py:37-39 */
  double timestep;  /* dt */
  double xtarget[4]; /* always zero in the reference: iterative_ilqr/utils/base.py:374 */
  double goal_tol;   /* 0.8, Simulator.sim :715 */
} i2lqr_params;

/* Obstacle (iterative_ilqr/utils/base.py:23-35).  present == 0 stands for `obstacle is None`.
 * moving_option: 0 = static (None in the reference), 1 = moving up, 2 = moving left. */
typedef struct i2lqr_obstacle {
  double x, y, width, height, spd;
  int32_t moving_option;
  int32_t present;
} i2lqr_obstacle;

typedef struct i2lqr_handle i2lqr_handle;

/* ---- library ---------------------------------------------------------------------------------------- */
int i2lqr_abi_version(void);
const char* i2lqr_last_error(void);
/* Fills p with the reference's defaults (iLqrParam(), KineticBicycleParam()). */
void i2lqr_default_params(i2lqr_params* p);

/* One handle = one GPU, one stream, `n_episodes` independent controllers (safe sets), each keeping its
 * most recent `num_ss_iter` laps of at most `lap_capacity` points.  Replaces iLqr.__init__
 * (iterative_ilqr/utils/base.py:305-330). */
int i2lqr_create(const i2lqr_params* p, int device, int n_episodes, int lap_capacity, i2lqr_handle** out);
int i2lqr_destroy(i2lqr_handle* h);
/* Re-reads the parameter block (scripts mutate iLqrParam between laps). */
int i2lqr_set_params(i2lqr_handle* h, const i2lqr_params* p);

/* ---- flat batch of independent candidate solves ----------------------------------------------------- */
/* Replaces B calls of control.iterative_ilqr.ilqr (iterative_ilqr/control/iterative_ilqr.py:7-85), each
 * set up as calc_input does (zero uvar, xvar[:,0] = x0: iterative_ilqr/utils/base.py:405-410).
 *   x0, x_term : (B,4)   lamb_in : (B) or NULL (= params.lamb)
 *   obs : n_obs entries, n_obs in {0 (no obstacle), 1 (shared), B}
 *   uvar_init : (B,2,N) warm-start inputs or NULL (= zeros, what calc_input passes)
 *   uvar : (B,2,N)  xvar : (B,4,N+1)  lamb_out : (B)  iters : (B) executed iLQR iterations
 * Any output pointer may be NULL.  Host-pointer form: H2D, solve, D2H inside the call. */
int i2lqr_solve_batch(i2lqr_handle* h, int64_t B, const double* x0, const double* x_term, const double* lamb_in,
                      const i2lqr_obstacle* obs, int64_t n_obs, const double* uvar_init, double* uvar, double* xvar,
                      double* lamb_out, int32_t* iters);
/* Same, all pointers resident on the handle's device; runs on `cuda_stream` (NULL = the handle's stream)
 * and does not synchronise the host.  ONE HANDLE PER STREAM: a handle owns a single set of device scratch
 * (parked candidates, long-horizon slab, statistics words), so consecutive calls on the same handle are
 * serialised on the device even when they are given different streams (each call waits for the previous
 * call's end event).  Batches that should overlap go through separate handles. */
int i2lqr_solve_batch_device(i2lqr_handle* h, int64_t B, const double* x0, const double* x_term,
                             const double* lamb_in, const i2lqr_obstacle* obs, int64_t n_obs, const double* uvar_init,
                             double* uvar, double* xvar, double* lamb_out, int32_t* iters, void* cuda_stream);

/* ---- safe set --------------------------------------------------------------------------------------- */
/* Replaces iLqr.add_trajectory (iterative_ilqr/utils/base.py:343-369): x is (T,4); Qfun[j] = T-1-j is
 * implied.  Requires T > num_horizon (as :348 does) and T <= lap_capacity. */
int i2lqr_add_lap(i2lqr_handle* h, int episode, const double* x, int T);
int i2lqr_num_laps(i2lqr_handle* h, int episode, int* n_laps);
/* `ilqr.obstacle = ...` (iterative_ilqr/result/ilqr_test_add_moving_obstacle.py:73-75); NULL = None. */
int i2lqr_set_obstacle(i2lqr_handle* h, int episode, const i2lqr_obstacle* obs);
/* EXTENSION: several obstacles per episode.  The reference's cost loop takes ONE Obstacle
 * (iterative_ilqr/control/ilqr_helper.py:16-24,32-51); a list of n <= 4 is priced as the SUM of its members' barriers, each
 * evaluated exactly as the single one.  list[0] is what i2lqr_set_obstacle would store; n = 0 or 1 is the reference's
 * behaviour.  Controller steps and batched episodes only (FP64, horizons whose step working set fits shared memory). */
int i2lqr_set_obstacle_list(i2lqr_handle* h, int episode, const i2lqr_obstacle* list, int n);
/* Batched forms for many-episode handles (one host-to-device copy instead of one per episode):
 * the same lap appended to EVERY episode's safe set (what each reference script does with the scripted first lap,
 * iterative_ilqr/tests/ilqr_test.py:49), and the obstacles of episodes [first, first+n) in one call
 * (obs == NULL clears them). */
int i2lqr_add_lap_all(i2lqr_handle* h, const double* x, int T);
int i2lqr_set_obstacles(i2lqr_handle* h, int first, int n, const i2lqr_obstacle* obs);
/* Replaces iLqr.select_close_ss (iterative_ilqr/utils/base.py:332-341): 1-norm K-nearest points of stored
 * lap `lap` (absolute lap number, as the reference's `iter` argument) in ascending (distance, index) order.
 * Writes min(K, T) indices and that count. */
int i2lqr_select_knn(i2lqr_handle* h, int episode, int lap, const double* x0, int K, int64_t* idx_out,
                     int32_t* n_out);

/* ---- controller step -------------------------------------------------------------------------------- */
/* Replaces the solve branch of iLqr.calc_input (iterative_ilqr/utils/base.py:384-478) for episodes
 * 0..n_episodes-1: 3 outer iterations x active laps x K candidates (lambda carried along each lap's
 * candidates), relax-cost ladder, lexicographic winner.
 *   x : (E,4) current states          u : (E,2)   x_pred : (E,4,N+1)   u_pred : (E,2,N)
 *   winner_lap : absolute lap number   winner_rank : position in the K-nearest order
 *   winner_idx : index of the winning point in its lap   winner_cost : its relaxed cost
 *   shrink : 1 when the winner is its lap's last point (horizon shrink / playback trigger, :472-478)
 * Any output pointer may be NULL. */
int i2lqr_solve_step(i2lqr_handle* h, int n_episodes, const double* x, double* u, double* x_pred, double* u_pred,
                     int32_t* winner_lap, int32_t* winner_rank, int64_t* winner_idx, double* winner_cost,
                     int32_t* shrink);
/* Per-candidate record of the last i2lqr_solve_step, shaped (E, 3, L, K): selected indices (-1 = unused
 * slot), relaxed costs, lambda after each candidate and executed iLQR iterations. */
int i2lqr_get_step_trace(i2lqr_handle* h, int n_episodes, int64_t* sel_idx, double* cand_cost, double* cand_lamb,
                         int32_t* cand_iters);

/* ---- controller step whose laps are split over several handles / GPUs --------------------------------- */
/* The K candidates of one lap form a sequential lambda chain, but the chains of different laps are independent
 * (iterative_ilqr/utils/base.py:386-455: `for id in range(min_iter, iter)`), so the active laps of ONE step can
 * be solved by different handles -- normally one per GPU, every handle holding the same safe set.  Handle
 * `part` of `n_parts` owns the active laps whose position in the lap loop satisfies pos % n_parts == part.
 * Per outer iteration `it` = 0..2 (the loop at :384):
 *     i2lqr_split_solve   K-nearest for every lap, the owned laps' chains, then one packed row per
 *                         (episode, lap) into rows_device (n_episodes, L, row_doubles) -- rows of laps this
 *                         handle does not own are zero;
 *     the caller all-gathers rows_device of all parts into all_rows_device (n_parts, n_episodes, L, row_doubles)
 *                         (ncclAllGather over NVLink; any copy that produces that layout will do);
 *     i2lqr_split_merge   unpacks the other parts' rows and applies the winner rule (:462-478); every handle then
 *                         holds the same x_pred / u_pred, which seed the next outer iteration (:398-401).
 * A row carries the lap's relaxed costs, lambdas and iteration counts (K each) and the trajectory of the lap's
 * first arg-min candidate, the only one of the lap that can win: row_doubles = 1 + 3K + 4(N+1) + 2N.
 * All device work is queued on `cuda_stream` (NULL = the handle's stream) without synchronising, so the calls
 * can be interleaved with NCCL on the same stream; i2lqr_split_end downloads the results (same arguments as
 * i2lqr_solve_step, any of which may be NULL) and blocks.  Results are bit-identical to i2lqr_solve_step. */
int i2lqr_split_row_doubles(i2lqr_handle* h, int64_t* row_doubles);
int i2lqr_split_begin(i2lqr_handle* h, int n_episodes, const double* x, void* cuda_stream);
int i2lqr_split_solve(i2lqr_handle* h, int n_episodes, int it, int part, int n_parts, double* rows_device,
                      void* cuda_stream);
int i2lqr_split_merge(i2lqr_handle* h, int n_episodes, int it, int part, int n_parts, const double* all_rows_device,
                      void* cuda_stream);
int i2lqr_split_end(i2lqr_handle* h, int n_episodes, double* u, double* x_pred, double* u_pred, int32_t* winner_lap,
                    int32_t* winner_rank, int64_t* winner_idx, double* winner_cost, int32_t* shrink, void* cuda_stream);

/* ---- batched closed-loop episodes (device-resident plant + controller) ------------------------------ */
/* Replaces Simulator.sim + KineticBicycle.forward_one_step (iterative_ilqr/utils/base.py:157-213,708-726)
 * and the lap post-processing of iterative_ilqr/tests/ilqr_test.py:56-60 for all episodes of the handle.
 * x_init : (E,4) state the first lap starts from (later laps start from 0, as update_memory_post_iter
 * resets the plant, :195); goal : (4) the initial lap's last row. */
int i2lqr_episodes_begin(i2lqr_handle* h, const double* x_init, const double* goal);
/* Runs one lap for every episode (at most max_sim_steps controller steps each); on return lap_len[e] is the
 * number of steps taken and the finished lap (last row snapped to the goal) is already appended to the
 * episode's safe set.  lap_x : optional (E, max_sim_steps+1, 4); lap_u : optional (E, max_sim_steps, 2). */
int i2lqr_episodes_run_lap(i2lqr_handle* h, int32_t* lap_len, double* lap_x, double* lap_u);

/* Plant noise of KineticBicycle.forward_dynamics (iterative_ilqr/utils/base.py:207-211): clipped Gaussians on
 * speed and heading, drawn on the device from a counter-based generator keyed by (seed, episode, lap, step).
 * seed = 0 (default) switches the noise off, as `set_zero_noise()` does in every reference script. */
int i2lqr_episodes_set_noise(i2lqr_handle* h, uint64_t seed);

/* Per-episode status of the last lap: 0 reached the goal, 1 stopped at max_sim_steps ("not finished",
 * :722-726), 2 playback ran out of inputs (IndexError in the reference), 3 lap too short/long to store. */
int i2lqr_episodes_status(i2lqr_handle* h, int32_t* status);

/* ---- the functions an iLQR iteration is made of, one by one ---------------------------------------------- */
/* The solve entry points above run these fused.  They are also exported separately, batched over B independent
 * problems, for callers that drive their own outer loop and for parity at function granularity; each runs the
 * same device code the solve kernels inline.  Arrays use the reference's shapes with a leading batch axis and
 * time on the LAST axis; host pointers; FP64 only.  obs / n_obs as in i2lqr_solve_batch.
 *
 * kinetic_bicycle(x, u, dt) (iterative_ilqr/systems/kinetic_bicycle.py:10-27): x (B,4), u (B,2) -> x_next (B,4). */
int i2lqr_dynamics_batch(i2lqr_handle* h, int64_t B, const double* x, const double* u, double* x_next);
/* get_A_matrix(velocity_vals, theta, acceleration_vals, num_horizon, dt) and get_B_matrix(theta, num_horizon, dt)
 * (iterative_ilqr/systems/kinetic_bicycle.py:30-52): v, theta, accel (B,N) -> A (B,4,4,N), Bm (B,4,2,N).  Either
 * output may be NULL; v and accel are only read for A. */
int i2lqr_linearize_batch(i2lqr_handle* h, int64_t B, const double* v, const double* theta, const double* accel, double* A,
                          double* Bm);
/* get_cost_derivation(ctrl_U, dX, ilqr_param, num_horizon, xvar, obstacle, sys_param)
 * (iterative_ilqr/control/ilqr_helper.py:9-56, with add_control_constraint :83-103 and repelling_cost_function
 * :59-64): uvar (B,2,N), dX (B,4,N+1), xvar (B,4,N+1) -> l_u (B,2,N), l_uu (B,2,2,N), l_x (B,4,N), l_xx (B,4,4,N). */
int i2lqr_cost_derivatives_batch(i2lqr_handle* h, int64_t B, const double* uvar, const double* dX, const double* xvar,
                                 const i2lqr_obstacle* obs, int64_t n_obs, double* l_u, double* l_uu, double* l_x,
                                 double* l_xx);
/* get_cost_final(x, x_terminal, matrix_Qterminal, obstacle, ilqr_param) (iterative_ilqr/control/ilqr_helper.py:106-150):
 * xvar (B,4,N+1), x_term (B,4) -> l_x (B,4), l_xx (B,4,4); the obstacle is advanced by params.obstacle_horizon. */
int i2lqr_cost_final_batch(i2lqr_handle* h, int64_t B, const double* xvar, const double* x_term, const i2lqr_obstacle* obs,
                           int64_t n_obs, double* l_x, double* l_xx);
/* backward_pass(xvar, uvar, x_terminal, dX, lamb, ...) (iterative_ilqr/control/iterative_ilqr.py:88-130):
 * -> k (B,2,N), K (B,2,4,N).  dX is taken as xvar - params.xtarget, which is what ilqr() maintains (:41-42);
 * lamb (B) or NULL (= params.lamb). */
int i2lqr_backward_pass_batch(i2lqr_handle* h, int64_t B, const double* xvar, const double* uvar, const double* x_term,
                              const double* lamb, const i2lqr_obstacle* obs, int64_t n_obs, double* k, double* K);
/* forward_pass(xvar, uvar, x_terminal, ..., matrix_k, matrix_K, sys_param) (iterative_ilqr/control/iterative_ilqr.py:133-160):
 * -> xvar_new (B,4,N+1), uvar_new (B,2,N), cost_new (B); any output may be NULL. */
int i2lqr_forward_pass_batch(i2lqr_handle* h, int64_t B, const double* xvar, const double* uvar, const double* x_term,
                             const double* k, const double* K, double* xvar_new, double* uvar_new, double* cost_new);

/* ---- instrumentation -------------------------------------------------------------------------------- */
/* Controller steps with few lambda chains (one controller, a few episodes) run as ONE launch in which idle SMs solve
 * the chains' candidates ahead of time for the most likely incoming lambda values; the chain takes a precomputed result
 * only when its lambda matches bit for bit, so results are identical to the sequential chain
 * (iterative_ilqr/utils/base.py:391-426).  Cumulative counters since i2lqr_create:
 *   [0] candidates served from a finished speculative solve  [1] served after waiting for one in flight
 *   [2] in the speculation set but not started: solved in place  [3] lambda not in the set: solved in place
 *   [4] speculative solves finished  [5] abandoned  [6] one-launch steps  [7] reserved */
int i2lqr_spec_stats(i2lqr_handle* h, int64_t* counters8);
/* Developer aid (environment I2LQR_SPEC_TIMELINE=1): where the last one-launch step spent its time, 8 words per (outer
 * iteration, chain) -- see csrc/i2lqr_api.cu.  *c_max = chains per outer iteration in the table (0: not recorded). */
int i2lqr_spec_timeline(i2lqr_handle* h, uint64_t* words, int64_t capacity_words, int32_t* c_max);

/* Device time (CUDA events on the stream of the call) of the last compute call, the number of kernels it
 * launched, and the iLQR iterations / solves it executed (device-side counters, for the roofline).  For
 * i2lqr_solve_batch the time spans the whole chunked pipeline (copies included); for the *_device form and
 * the controller step it spans the kernels only. */
int i2lqr_last_stats(i2lqr_handle* h, double* kernel_ms, int64_t* launches, int64_t* ilqr_iterations,
                     int64_t* solves);

/* Refreshes the statistics after an asynchronous *_device call; synchronises `cuda_stream`. */
int i2lqr_sync_stats(i2lqr_handle* h, void* cuda_stream);
/* Dependent-free DFMA micro-benchmark: the measured FP64 FMA-pipe peak of `device` in TFLOP/s (best of
 * `reps` launches), the roofline denominator for this path (MEASURED_PEAKS.json has no FP64 figure). */
int i2lqr_measure_fp64_peak(int device, int reps, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* I2LQR_H_ */
