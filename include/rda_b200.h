/* rda_b200.h — C ABI of the B200-native RDA ADMM-MPC hot path.
 *
 * The reference (hanruihua/RDA-planner) has no FFI: its boundary is the Python class
 * RDA_planner/rda_solver.py::RDA_solver (constructor :18-22, iterative_solve :573-610,
 * assign_adjust_parameter :426-434, get_adjust_parameter :1055-1056, reset :1060-1068).
 * Each entry point below replaces the reference code cited next to it.  All pointers in
 * rda_inputs / rda_outputs are DEVICE pointers to float32 arrays owned by the caller
 * (PyTorch); the handle owns only the persistent warm-start state the reference keeps in
 * cvxpy Parameters (rda_solver.py:112-182).  No entry point throws, allocates host memory
 * per call, synchronises the device (except create/destroy/export) or falls back to the
 * CPU.  Return value: 0 ok, <0 usage error (RDA_E_*), >0 a cudaError_t.
 */
#ifndef RDA_B200_H
#define RDA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDA_MAX_EDGE 8        /* max_edge_num  (obstacle rows)  */
#define RDA_MAX_ROBOT_EDGE 8  /* rows of car_tuple.G            */

enum { RDA_DYN_ACKER = 0, RDA_DYN_DIFF = 1, RDA_DYN_OMNI = 2 };       /* rda_solver.py:446-451 */
enum { RDA_OBS_POLYGON = 0, RDA_OBS_CIRCLE = 1 };                     /* cone flag :158,:512   */
enum { RDA_ROBOT_POLYGON = 0, RDA_ROBOT_DISC = 1 };                   /* car_tuple.cone_type 'Rpositive' / 'norm2', :1034-1039 */
enum { RDA_E_ARG = -1, RDA_E_UNSUPPORTED = -2, RDA_E_NOMEM = -3 };

/* per-instance status bits (device word), mirroring the reference's keep-previous-iterate
 * rules (rda_solver.py:696-700, :791-793) */
enum { RDA_ST_SU_NOT_CONVERGED = 1, RDA_ST_SU_NONFINITE = 2, RDA_ST_CELL_FALLBACK = 4,
       RDA_ST_EARLY_STOP = 8 };

typedef struct rda_config {
  int batch;          /* B planning instances sharing this configuration (reference: 1)  */
  int receding;       /* T   rda_solver.py:34                                            */
  int max_obs_num;    /* N   :38                                                         */
  int max_edge_num;   /* E   :39   (<= RDA_MAX_EDGE)                                      */
  int robot_edges;    /* R = G.shape[0] (<= RDA_MAX_ROBOT_EDGE); 3 for a disc body          */
  int dynamics;       /* RDA_DYN_*  :40                                                  */
  int accelerated;    /* :47                                                             */
  int su_fp64;        /* 1: su-QP interior point arithmetic in float64 (default), 0: float32 */
  float step_time;    /* dt  :43                                                         */
  float wheelbase;    /* L   :36                                                         */
  float max_speed[2]; /* :37                                                             */
  float acce_bound[2];/* max_acce * dt  :44                                              */
  float ws, wu;       /* :218-219 (fixed at construction, as in the reference)           */
  float G[RDA_MAX_ROBOT_EDGE * 2]; /* robot half-spaces, rows CCW (car_tuple.G)          */
  float h[RDA_MAX_ROBOT_EDGE];     /* car_tuple.h                                        */
  int robot_cone;     /* RDA_ROBOT_*: polygon body (rows of G counter-clockwise) or the disc body
                         G = [[1,0],[0,1],[0,0]], h = (cx, cy, -r) of cone_type 'norm2' (:1034-1039) */
} rda_config;

typedef struct rda_tunables {   /* rda_solver.py:185-201, :426-434 */
  float slack_gain, max_sd, min_sd, ro1, ro2;
  float z_theta;        /* tie-break of the slack z in [0, stuff]; 0.5 default */
} rda_tunables;

/* Inputs of one batched solve.  Layouts (row-major, last index fastest):
 *   nom_s  [B][3][T+1]   nominal states   (iterative_solve arg nom_s,  :573,:584)
 *   nom_u  [B][2][T]     nominal controls (arg nom_u)
 *   ref_s  [B][3][T+1]   reference states (arg ref_states, :580)
 *   ref_speed [B]        (:581)
 *   obs_A  [B][N][Tc][E][2], obs_b [B][N][Tc][E]   half-spaces per obstacle slot and copy
 *          (assign_obstacle_parameter :483-526; rows zero-padded; Tc = T+1 if
 *          obs_time_varying else 1)
 *   obs_kind [B][N] int32 RDA_OBS_*;  obs_count [B] int32 = len(obstacle_list) before padding
 *          (0 => LamMuZ/xi/zeta skipped for that instance, :625)                              */
typedef struct rda_inputs {
  const float *nom_s, *nom_u, *ref_s, *ref_speed;
  const float *obs_A, *obs_b;
  const int32_t *obs_kind, *obs_count;
  int obs_time_varying;
} rda_inputs;

/* Outputs (device): u_opt [B][2][T], s_opt [B][3][T+1] (solution of the last su-QP, :603-610),
 * resi_pri [B], resi_dual [B] (:604-608), status [B] int32 (RDA_ST_*), iters [B] int32.     */
typedef struct rda_outputs {
  float *u_opt, *s_opt, *resi_pri, *resi_dual;
  int32_t *status, *iters;
} rda_outputs;

typedef struct rda_handle rda_handle;

/* RDA_solver.__init__ / definition() (:18-78, :112-201): allocate persistent state, set the
 * reference initial values (all zero, para_dis = 1). */
int rda_create(const rda_config *cfg, const rda_tunables *tun, rda_handle **out);
int rda_destroy(rda_handle *h);
/* assign_adjust_parameter (:426-434) / get_adjust_parameter (:1055-1056) */
int rda_set_tunables(rda_handle *h, const rda_tunables *tun);
int rda_get_tunables(const rda_handle *h, rda_tunables *tun);
/* RDA_solver.reset (:1060-1068): clears lam'A and lam'b only. */
int rda_reset(rda_handle *h, void *cuda_stream);
/* Clear ALL warm-start state back to the constructor values (extension; used by benchmarks). */
int rda_cold_start(rda_handle *h, void *cuda_stream);
/* iterative_solve (:573-610): iter_num ADMM iterations (early stop per instance when both
 * residuals < iter_threshold, :594), enqueued on cuda_stream, no host synchronisation. */
int rda_solve(rda_handle *h, const rda_inputs *in, const rda_outputs *out, int iter_num,
              float iter_threshold, void *cuda_stream);
/* Single phases, for unit tests and profiling: begin (load nominal + obstacles), one su-QP
 * (su_prob_solve :692-700 + assign_state_parameter :436-460), one LamMuZ + multiplier update
 * (:702-741, :529-542, :639-690), and the output stage. */
int rda_begin(rda_handle *h, const rda_inputs *in, float iter_threshold, void *cuda_stream);
int rda_step_su(rda_handle *h, void *cuda_stream);
int rda_step_lammuz(rda_handle *h, void *cuda_stream);
int rda_finish(rda_handle *h, const rda_outputs *out, void *cuda_stream);

/* Persistent buffers (device pointers, float32 unless noted), for tests / checkpointing:
 * ids below; returns element count in *count. */
enum { RDA_BUF_LAM = 0,   /* [B][N][E][T]  columns 1..T of para_lam (:141)   */
       RDA_BUF_MU = 1,    /* [B][N][R][T]                        (:142)      */
       RDA_BUF_Z = 2,     /* [B][N][T]                           (:143)      */
       RDA_BUF_XI = 3,    /* [B][2][N][T]  rows 1..T of para_xi  (:144)      */
       RDA_BUF_ZETA = 4,  /* [B][N][T]                           (:145)      */
       RDA_BUF_DIS = 5,   /* [B][T]        para_dis              (:119)      */
       RDA_BUF_COEF = 6,  /* [B][5][N][T]  su-QP hinge inputs: a_x, a_y (lam'A :172), c0, g_x, g_y */
       RDA_BUF_PREF = 7,  /* [B][2][T]     positions c0 refers to            */
       RDA_BUF_CUR_S = 8, /* [B][3][T+1]   current nominal (para_s :117)     */
       RDA_BUF_CUR_U = 9, /* [B][2][T]     para_u (:118)                     */
       RDA_BUF_COUNTERS = 10 /* int32 [8]: cells fast, cells slow, cells fallback, su iterations ... */
};
int rda_get_buffer(rda_handle *h, int id, void **dev_ptr, size_t *count);
/* Copy a persistent buffer to (to_handle = 0) or from (to_handle = 1) caller-owned device
 * memory of the same element count, asynchronously on cuda_stream (checkpoint / resume). */
int rda_copy_buffer(rda_handle *h, int id, void *user_dev_ptr, int to_handle, void *cuda_stream);

/* number of kernels the last rda_solve / phase call enqueued (for bench.py's gpu_launches) */
int rda_last_launch_count(const rda_handle *h);
const char *rda_version(void);

/* ------------------------------------------------------------------------------------------
 * Front end: the steps either side of the solve, batched and stateless.
 * All pointers are DEVICE pointers; every call only enqueues one kernel on cuda_stream.
 * ------------------------------------------------------------------------------------------ */
#define RDA_MAX_SHAPES 64     /* raw obstacles per instance handed to rda_convert_obstacles */

/* MPC.pre_process (mpc.py:251-291) with closest_point :338-353, inter_point :355-383,
 * range_cir_seg :385-423, wraptopi :431-438 and motion_predict_model_* :293-336, for B
 * robots on ONE reference path.
 *   state [B][3]; cur_vel [B][2][T] (the controls of the previous step, MPC.cur_vel_array);
 *   ref_speed [B] (already multiplied by the gear); path [P][3] (x, y, heading);
 *   start_index [B] (MPC.cur_index; NULL = 0); threshold / ind_range: closest_point kwargs
 *   (0.1 / 10).  Out: nom_s, ref_s [B][3][T+1] (the solver's nom_s / ref_s), near_index [B]
 *   (the new MPC.cur_index).                                                              */
int rda_pre_process(int B, int T, int dynamics, float dt, float wheelbase, const float *state,
                    const float *cur_vel, const float *ref_speed, const float *path, int P,
                    const int32_t *start_index, float threshold, int ind_range, float *nom_s,
                    float *ref_s, int32_t *near_index, void *cuda_stream);

/* MPC.convert_rda_obstacle (mpc.py:189-218: conversion + optional distance sort) with
 * convert_inequal_circle :440-458, convert_inequal_polygon :460-474, gen_inequal_global
 * :476-510, is_convex_and_ordered :518-549, followed by RDA_solver.assign_obstacle_parameter
 * (rda_solver.py:483-526: keep the first N, pad a short list by repeating its last element,
 * zero rows beyond the shape's own).  Up to M <= RDA_MAX_SHAPES raw shapes per instance:
 *   shape_kind [B][M] RDA_OBS_*; shape_nv [B][M] vertices of a polygon (3..E);
 *   shape_xy [B][M][RDA_MAX_EDGE][2] polygon vertices, or the disc centre in entry 0;
 *   shape_radius [B][M]; shape_vel [B][M][2]; shape_count [B]; state [B][3] (sort key origin,
 *   may be NULL when order = 0).  time_varying selects the output layout (T+1 copies per
 *   obstacle, moved by velocity * t * dt when |velocity| > 0.01, or one copy at t = 0).
 * Out: obs_A, obs_b, obs_kind, obs_count exactly as rda_inputs expects them.              */
int rda_convert_obstacles(int B, int M, int N, int T, int E, float dt, int time_varying, int order,
                          const float *state, const int32_t *shape_kind, const int32_t *shape_nv,
                          const float *shape_xy, const float *shape_radius, const float *shape_vel,
                          const int32_t *shape_count, float *obs_A, float *obs_b, int32_t *obs_kind,
                          int32_t *obs_count, void *cuda_stream);

/* Arrive rule of MPC.control (mpc.py:170-185, single gear): instances whose near_index >=
 * P - goal_index_threshold get u_opt = 0 and arrive = 1; cur_vel (may be NULL) receives the
 * controls kept as the next step's nominal (mpc.py:186).  u_opt, cur_vel [B][2][T].        */
int rda_post_process(int B, int T, int P, int goal_index_threshold, const int32_t *near_index,
                     float *u_opt, float *cur_vel, int32_t *arrive, void *cuda_stream);

/* Gear changes (enable_reverse, mpc.py:139-144, :166-183, split_path :232-249).  The reference path is cut into
 * n_curves single-gear curves, curve c = waypoints [curve_start[c], curve_start[c+1]) of `path`; every robot follows
 * curve curve_index[b].  rda_pre_process_curves is rda_pre_process on the robot's current curve (near_index is
 * relative to the curve).  rda_post_process_gear applies mpc.py:166-185: at the end of a curve the robot switches
 * to the next one (cur_index back to 0, controls kept); past the last curve the controls are zeroed and arrive is
 * set (curve_index then stays on the last curve: the reference would raise IndexError on the next call).
 * gear [B] (output) is the gear flag (+1 / -1) of the robot's curve BEFORE the update; the caller multiplies the
 * solver's reference speed with it (mpc.py:161). */
int rda_pre_process_curves(int B, int T, int dynamics, float dt, float wheelbase, const float *state,
                           const float *cur_vel, const float *ref_speed, const float *path, int n_curves,
                           const int32_t *curve_start, const int32_t *curve_index, const int32_t *start_index,
                           float threshold, int ind_range, float *nom_s, float *ref_s, int32_t *near_index,
                           void *cuda_stream);
int rda_post_process_gear(int B, int T, int n_curves, const int32_t *curve_start, int goal_index_threshold,
                          int32_t *near_index, int32_t *curve_index, float *u_opt, float *cur_vel,
                          int32_t *arrive, void *cuda_stream);

/* state [B][3] advanced in place by one step of the nonlinear model with the first control of
 * u_opt [B][2][T] (mpc.py:293-336; what the examples' simulator does between control calls). */
int rda_motion_predict(int B, int T, int dynamics, float dt, float wheelbase, const float *u_opt,
                       float *state, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
