// rda_frontend.cu — batched kernels for the steps either side of the solver:
// reference pre-processing, obstacle conversion, the arrive rule and the state advance of the closed
// loop.  Stateless C ABI (include/rda_b200.h, "front end" section); the per-instance arithmetic lives in
// frontend.cuh, shared with the CPU tests.  These passes are tiny next to the solve (one thread or one
// small CTA per instance, a few hundred bytes each); they exist so that a closed-loop step never leaves
// the device.
#include <cuda_runtime.h>
#include "../../include/rda_b200.h"
#include "rda_hd.h"
#include "frontend.cuh"

using namespace rda;

#define RDA_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

namespace {

__global__ void k_pre_process(int B, int T, int dynamics, float dt, float L, const float* state, const float* cur_vel,
                              const float* ref_speed, const float* path, int P, const int* start_index,
                              float threshold, int ind_range, float* nom_s, float* ref_s, int* near_index) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int near = pre_process_one(dynamics, T, (double)dt, (double)L, state + 3 * (size_t)b,
                                   cur_vel + (size_t)b * 2 * T, (double)ref_speed[b], path, P,
                                   start_index ? start_index[b] : 0, (double)threshold, ind_range,
                                   nom_s + (size_t)b * 3 * (T + 1), ref_s + (size_t)b * 3 * (T + 1));
  near_index[b] = near;
}

// the same on the single-gear curve the robot currently follows (enable_reverse, mpc.py:139-144)
__global__ void k_pre_process_curves(int B, int T, int dynamics, float dt, float L, const float* state, const float* cur_vel,
                                     const float* ref_speed, const float* path, int n_curves, const int* curve_start,
                                     const int* curve_index, const int* start_index, float threshold, int ind_range,
                                     float* nom_s, float* ref_s, int* near_index) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int c = curve_index[b];
  c = c < 0 ? 0 : (c >= n_curves ? n_curves - 1 : c);
  const int p0 = curve_start[c], len = curve_start[c + 1] - p0;
  const int near = pre_process_one(dynamics, T, (double)dt, (double)L, state + 3 * (size_t)b,
                                   cur_vel + (size_t)b * 2 * T, (double)ref_speed[b], path + 3 * (size_t)p0, len,
                                   start_index ? start_index[b] : 0, (double)threshold, ind_range,
                                   nom_s + (size_t)b * 3 * (T + 1), ref_s + (size_t)b * 3 * (T + 1));
  near_index[b] = near;
}

// one CTA per instance: keys -> stable ranks -> rows of the N slots
__global__ void __launch_bounds__(RDA_MAX_SHAPES)
k_convert_obstacles(int B, int M, int N, int T, int E, float dt, int time_varying, int order, const float* state,
                    const int* shape_kind, const int* shape_nv, const float* shape_xy, const float* shape_radius,
                    const float* shape_vel, const int* shape_count, float* obs_A, float* obs_b, int* obs_kind,
                    int* obs_count) {
  __shared__ double keys[RDA_MAX_SHAPES];
  __shared__ int sorted[RDA_MAX_SHAPES];
  const int b = blockIdx.x;
  if (b >= B) return;
  const int j = threadIdx.x;
  int count = shape_count[b];
  if (count > M) count = M;
  if (count < 0) count = 0;
  const size_t sb = (size_t)b * M;
  if (j < count) keys[j] = order ? obstacle_key(shape_kind[sb + j], shape_nv[sb + j], shape_xy + (sb + j) * RDA_MAX_EDGE * 2,
                                               (double)state[3 * b], (double)state[3 * b + 1])
                                 : (double)j;
  __syncthreads();
  if (j < count) {
    int before = 0;
    for (int i = 0; i < count; ++i)
      if (keys[i] < keys[j] || (keys[i] == keys[j] && i < j)) ++before;
    sorted[before] = j;
  }
  __syncthreads();
  if (j == 0) obs_count[b] = count;
  const int Tc = time_varying ? T + 1 : 1;
  for (int n = j; n < N; n += blockDim.x) {
    float* A = obs_A + ((size_t)b * N + n) * Tc * E * 2;
    float* bb = obs_b + ((size_t)b * N + n) * Tc * E;
    if (count == 0) {
      for (int i = 0; i < Tc * E; ++i) { A[2 * i] = 0.f; A[2 * i + 1] = 0.f; bb[i] = 0.f; }
      obs_kind[(size_t)b * N + n] = RDA_OBS_POLYGON;
      continue;
    }
    const int src = sorted[n < count ? n : count - 1];
    const int kind = shape_kind[sb + src], nv = shape_nv[sb + src];
    const float* xy = shape_xy + (sb + src) * RDA_MAX_EDGE * 2;
    const double rad = shape_radius[sb + src];
    const double vx = shape_vel[(sb + src) * 2], vy = shape_vel[(sb + src) * 2 + 1];
    obs_kind[(size_t)b * N + n] = kind;
    for (int t = 0; t < Tc; ++t) obstacle_rows(kind, nv, xy, rad, vx, vy, t, (double)dt, E, A + (size_t)t * E * 2, bb + (size_t)t * E);
  }
}

// arrive rule (mpc.py:170-185 without gear changes) and the controls kept as next nominal (mpc.py:186)
__global__ void k_post_process(int B, int T, int P, int goal_index_threshold, const int* near_index, float* u_opt,
                               float* cur_vel, int* arrive) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 2 * T) return;
  const int b = i / (2 * T);
  const bool arr = near_index[b] >= P - goal_index_threshold;
  float u = u_opt[i];
  if (arr) { u = 0.f; u_opt[i] = 0.f; }
  if (cur_vel) cur_vel[i] = u;
  if (arrive && i == b * 2 * T) arrive[b] = arr ? 1 : 0;
}

// end-of-curve rule with gear changes (mpc.py:166-185), one thread per robot
__global__ void k_post_process_gear(int B, int T, int n_curves, const int* curve_start, int goal_index_threshold,
                                    int* near_index, int* curve_index, float* u_opt, float* cur_vel, int* arrive) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int c = curve_index[b];
  c = c < 0 ? 0 : (c >= n_curves ? n_curves - 1 : c);
  const int len = curve_start[c + 1] - curve_start[c];
  bool arr = false;
  if (near_index[b] >= len - goal_index_threshold) {
    if (c + 1 < n_curves) { curve_index[b] = c + 1; near_index[b] = 0; }     // next curve, controls kept
    else arr = true;                                                            // past the last curve
  }
  float* u = u_opt + (size_t)b * 2 * T;
  for (int i = 0; i < 2 * T; ++i) {
    if (arr) u[i] = 0.f;
    if (cur_vel) cur_vel[(size_t)b * 2 * T + i] = u[i];
  }
  if (arrive) arrive[b] = arr ? 1 : 0;
}

__global__ void k_motion_predict(int B, int T, int dynamics, float dt, float L, const float* u_opt, float* state) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double s[3] = {state[3 * b], state[3 * b + 1], state[3 * b + 2]};
  double o[3];
  motion_predict(dynamics, (double)dt, (double)L, s, (double)u_opt[(size_t)b * 2 * T], (double)u_opt[(size_t)b * 2 * T + T], o);
  state[3 * b] = (float)o[0]; state[3 * b + 1] = (float)o[1]; state[3 * b + 2] = (float)o[2];
}

}  // namespace

extern "C" {

int rda_pre_process(int B, int T, int dynamics, float dt, float wheelbase, const float* state, const float* cur_vel,
                    const float* ref_speed, const float* path, int P, const int32_t* start_index, float threshold,
                    int ind_range, float* nom_s, float* ref_s, int32_t* near_index, void* stream) {
  if (B < 1 || T < 1 || P < 1 || dynamics < 0 || dynamics > 2) return RDA_E_ARG;
  if (!state || !cur_vel || !ref_speed || !path || !nom_s || !ref_s || !near_index) return RDA_E_ARG;
  k_pre_process<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, T, dynamics, dt, wheelbase, state, cur_vel, ref_speed,
                                                                   path, P, start_index, threshold, ind_range, nom_s,
                                                                   ref_s, near_index);
  RDA_CUDA(cudaGetLastError());
  return 0;
}

int rda_pre_process_curves(int B, int T, int dynamics, float dt, float wheelbase, const float* state, const float* cur_vel,
                           const float* ref_speed, const float* path, int n_curves, const int32_t* curve_start,
                           const int32_t* curve_index, const int32_t* start_index, float threshold, int ind_range,
                           float* nom_s, float* ref_s, int32_t* near_index, void* stream) {
  if (B < 1 || T < 1 || n_curves < 1 || dynamics < 0 || dynamics > 2) return RDA_E_ARG;
  if (!state || !cur_vel || !ref_speed || !path || !curve_start || !curve_index || !nom_s || !ref_s || !near_index) return RDA_E_ARG;
  k_pre_process_curves<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, T, dynamics, dt, wheelbase, state, cur_vel,
                                                                          ref_speed, path, n_curves, curve_start, curve_index,
                                                                          start_index, threshold, ind_range, nom_s, ref_s,
                                                                          near_index);
  RDA_CUDA(cudaGetLastError());
  return 0;
}

int rda_post_process_gear(int B, int T, int n_curves, const int32_t* curve_start, int goal_index_threshold,
                          int32_t* near_index, int32_t* curve_index, float* u_opt, float* cur_vel, int32_t* arrive,
                          void* stream) {
  if (B < 1 || T < 1 || n_curves < 1 || !curve_start || !near_index || !curve_index || !u_opt) return RDA_E_ARG;
  k_post_process_gear<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, T, n_curves, curve_start, goal_index_threshold,
                                                                         near_index, curve_index, u_opt, cur_vel, arrive);
  RDA_CUDA(cudaGetLastError());
  return 0;
}

int rda_convert_obstacles(int B, int M, int N, int T, int E, float dt, int time_varying, int order, const float* state,
                          const int32_t* shape_kind, const int32_t* shape_nv, const float* shape_xy,
                          const float* shape_radius, const float* shape_vel, const int32_t* shape_count, float* obs_A,
                          float* obs_b, int32_t* obs_kind, int32_t* obs_count, void* stream) {
  if (B < 1 || N < 1 || T < 1 || M < 1) return RDA_E_ARG;
  if (M > RDA_MAX_SHAPES || E < 3 || E > RDA_MAX_EDGE) return RDA_E_UNSUPPORTED;
  if (!shape_kind || !shape_nv || !shape_xy || !shape_radius || !shape_vel || !shape_count) return RDA_E_ARG;
  if (!obs_A || !obs_b || !obs_kind || !obs_count || (order && !state)) return RDA_E_ARG;
  k_convert_obstacles<<<B, RDA_MAX_SHAPES, 0, (cudaStream_t)stream>>>(B, M, N, T, E, dt, time_varying, order, state, shape_kind,
                                                                     shape_nv, shape_xy, shape_radius, shape_vel,
                                                                     shape_count, obs_A, obs_b, obs_kind, obs_count);
  RDA_CUDA(cudaGetLastError());
  return 0;
}

int rda_post_process(int B, int T, int P, int goal_index_threshold, const int32_t* near_index, float* u_opt,
                     float* cur_vel, int32_t* arrive, void* stream) {
  if (B < 1 || T < 1 || !near_index || !u_opt) return RDA_E_ARG;
  const int n = B * 2 * T;
  k_post_process<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(B, T, P, goal_index_threshold, near_index, u_opt,
                                                                    cur_vel, arrive);
  RDA_CUDA(cudaGetLastError());
  return 0;
}

int rda_motion_predict(int B, int T, int dynamics, float dt, float wheelbase, const float* u_opt, float* state,
                       void* stream) {
  if (B < 1 || T < 1 || dynamics < 0 || dynamics > 2 || !u_opt || !state) return RDA_E_ARG;
  k_motion_predict<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, T, dynamics, dt, wheelbase, u_opt, state);
  RDA_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
