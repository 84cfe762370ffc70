// su_batched.cuh — the su-QP of a whole (sub-)batch as a pipeline of wide kernels.
//
// Same problem, same Mehrotra predictor-corrector interior point method and the same Riccati recursion
// as su_solver.cuh (reference lines there: rda_solver.py:216-231, :313-387, :692-700, :831-872,
// :911-947, :1011-1032), but mapped on the GPU the other way round.  k_su (one warp per instance) runs
// the serial Riccati recursion redundantly in all 32 lanes and keeps ~11 instances per SM resident; at
// large batches that leaves most of the machine idle.  Here every phase of an interior point iteration
// is its own kernel over the whole batch:
//
//   stage kernels (one thread per (stage t, instance b), b fastest => every access coalesced):
//     ksb_assemble  update of the previous iteration (slacks, multipliers, iterate) fused with the
//                   predictor assembly (stage gradient, barrier weights, hinge Hessian, d_t eliminated)
//     ksb_affine    step length and predicted complementarity of the affine direction
//     ksb_corrector corrector right-hand side (needs sigma*mu: one reduction per instance in between)
//     ksb_steplen   step length of the combined direction
//   recursion kernels (one thread per instance, operands of the next stage prefetched into registers):
//     ksb_factor    convergence test, Riccati factorisation + forward sweep (affine direction)
//     ksb_solve     Riccati solve with the stored factors + forward sweep (combined direction)
//
// All state lives in a global workspace laid out [index][instance]; per-instance reductions over the
// horizon go through per-stage partials summed in a fixed order by the next per-instance kernel
// (deterministic; the kernel boundary is the barrier).  Converged instances drop out (active flag);
// kernels return immediately once nobody is active.  tests/ run the same kernels on the CPU through
// a serial launch emulation (RDA_SB_EMULATE, oracle/cpu_port).
#pragma once
#include "su_solver.cuh"

namespace rda {

struct SuBatch {
  int nb, T, N;
  double *s[2], *u[2], *dd[2];      // ping-pong iterate: 3(T+1), 2T, T
  float *lins, *ref, *pref;         // 3(T+1), 3(T+1), 2T
  double *Aj, *Bj, *Cj, *Skk, *Sgk; // 2T, 6T, 3T, T, T
  double *bs, *bnu;                 // 10T
  float *hx, *hy, *hc;              // N*T  [o][t][b]
  double *hs, *hnu;                 // N*T
  double *Wm, *wb, *Ed, *g5q, *gw;  // 3T, 5T, 3T, T, 8T
  double *K, *Lc, *kf;              // 10T, 3T, 2T
  double *dza, *dva, *dz, *dv;      // 5(T+1), 3T, 5(T+1), 3T
  // per instance
  double *p_mu, *p_r, *p_step, *p_rmax, *p_s0, *p_s1, *p_s2;   // per-stage partial reductions, T each
  double *mu, *sigma_mu, *alpha, *vref;
  int *active, *st, *its;
  int *list;                        // compact, ascending list of the active instances (ksb_compact)
  int *n_active;                    // its length
};

// bytes of workspace for nb instances; if base != nullptr the pointers are set
inline size_t su_batch_layout(int nb, int T, int N, SuBatch* w, char* base) {
  size_t off = 0;
  auto take = [&](size_t n, size_t elt) {
    off = (off + 255) & ~(size_t)255;
    char* p = base ? base + off : nullptr;
    off += n * elt * (size_t)nb;
    return p;
  };
  SuBatch tmp;
  SuBatch& W = w ? *w : tmp;
  W.nb = nb; W.T = T; W.N = N;
  const size_t NT = (size_t)N * T;
  for (int k = 0; k < 2; ++k) {
    W.s[k] = (double*)take(3 * (T + 1), 8); W.u[k] = (double*)take(2 * T, 8); W.dd[k] = (double*)take(T, 8);
  }
  W.lins = (float*)take(3 * (T + 1), 4); W.ref = (float*)take(3 * (T + 1), 4); W.pref = (float*)take(2 * T, 4);
  W.Aj = (double*)take(2 * T, 8); W.Bj = (double*)take(6 * T, 8); W.Cj = (double*)take(3 * T, 8);
  W.Skk = (double*)take(T, 8); W.Sgk = (double*)take(T, 8);
  W.bs = (double*)take(10 * T, 8); W.bnu = (double*)take(10 * T, 8);
  W.hx = (float*)take(NT, 4); W.hy = (float*)take(NT, 4); W.hc = (float*)take(NT, 4);
  W.hs = (double*)take(NT, 8); W.hnu = (double*)take(NT, 8);
  W.Wm = (double*)take(3 * T, 8); W.wb = (double*)take(5 * T, 8); W.Ed = (double*)take(3 * T, 8);
  W.g5q = (double*)take(T, 8); W.gw = (double*)take(8 * T, 8);
  W.K = (double*)take(10 * T, 8); W.Lc = (double*)take(3 * T, 8); W.kf = (double*)take(2 * T, 8);
  W.dza = (double*)take(5 * (T + 1), 8); W.dva = (double*)take(3 * T, 8);
  W.dz = (double*)take(5 * (T + 1), 8); W.dv = (double*)take(3 * T, 8);
  W.p_mu = (double*)take(T, 8); W.p_r = (double*)take(T, 8); W.p_step = (double*)take(T, 8);
  W.p_rmax = (double*)take(T, 8); W.p_s0 = (double*)take(T, 8); W.p_s1 = (double*)take(T, 8); W.p_s2 = (double*)take(T, 8);
  W.mu = (double*)take(1, 8); W.sigma_mu = (double*)take(1, 8); W.alpha = (double*)take(1, 8); W.vref = (double*)take(1, 8);
  W.active = (int*)take(1, 4); W.st = (int*)take(1, 4); W.its = (int*)take(1, 4); W.list = (int*)take(1, 4);
  off = (off + 255) & ~(size_t)255;
  W.n_active = (int*)(base ? base + off : nullptr);
  off += 256;
  return off;
}

#if defined(__CUDACC__) || defined(RDA_SB_EMULATE)

#define SB_AT(arr, i) (arr)[(size_t)(i) * nb + b]

// one inequality row of stage t from register copies of (u_t, u_{t-1}, d_t); mirrors su_row
struct SbRow { double g, sgn; int comp; bool rate, live; };
__device__ __forceinline__ SbRow sb_row(const SuParams& P, int t, int c, const double* ut, const double* up, double dt_) {
  SbRow r;
  r.rate = c >= 6;
  r.live = true;
  const bool hi = (c & 1) == 0;
  r.sgn = hi ? -1.0 : 1.0;
  if (c < 4) {
    const int k = c >> 1;
    r.comp = 3 + k;
    const double uv = ut[k], m = P.umax[k];
    r.g = hi ? m - uv : uv + m;
  } else if (c < 6) {
    r.comp = 5;
    const double lo = P.dmin > 0 ? (double)P.dmin : 0.0;
    r.g = hi ? (double)P.dmax - dt_ : dt_ - lo;
    r.live = P.N > 0;
  } else {
    const int k = (c - 6) >> 1;
    r.comp = 3 + k;
    r.live = t >= 1;
    const double du = r.live ? ut[k] - up[k] : 0.0;
    r.g = hi ? (double)P.ab[k] - du : (double)P.ab[k] + du;
  }
  return r;
}

// directional derivative of a row along a step: dvt = (du0, du1, dd) of stage t, dup = (du0, du1) of stage t-1
__device__ __forceinline__ double sb_row_dir(const SbRow& r, const double* dvt, const double* dup) {
  double x = dvt[r.comp - 3];
  if (r.rate) x -= dup[r.comp - 3];
  return r.sgn * x;
}

// ---- setup: transposition of the inputs, linearisation, aggregated rotation terms ----------------
// cur_s/ref_s [b][3][T+1], cur_u [b][2][T], pref [b][2][T], coef [b][5][N][T], dis [b][T]
__global__ void ksb_setup(SuBatch W, SuParams P, const float* cur_s, const float* cur_u, const float* ref_s,
                          const float* pref, const float* coef, const float* dis, const float* ref_speed, const int* done) {
  const int nb = W.nb, T = W.T, N = W.N;
  const int b = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (b >= nb) return;
  if (t == 0) {
    const int act = done[b] ? 0 : 1;
    W.active[b] = act; W.st[b] = 1; W.its[b] = 0;
    W.mu[b] = 0; W.sigma_mu[b] = 0; W.alpha[b] = 0;
    W.vref[b] = ref_speed[b];
  }
  if (done[b]) return;
  const float* cs = cur_s + (size_t)b * 3 * (T + 1);
  const float* rf = ref_s + (size_t)b * 3 * (T + 1);
  const float* cu = cur_u + (size_t)b * 2 * T;
  for (int r = 0; r < 3; ++r) {
    SB_AT(W.lins, 3 * t + r) = cs[r * (T + 1) + t];
    SB_AT(W.ref, 3 * t + r) = rf[r * (T + 1) + t];
    if (t == T - 1) {
      SB_AT(W.lins, 3 * T + r) = cs[r * (T + 1) + T];
      SB_AT(W.ref, 3 * T + r) = rf[r * (T + 1) + T];
    }
  }
  const double st[3] = {(double)cs[t], (double)cs[(T + 1) + t], (double)cs[2 * (T + 1) + t]};
  const double ut[2] = {(double)cu[t], (double)cu[T + t]};
  double Aj[2], Bj[6], Cj[3];
  su_linearise<double>(P, st, ut, Aj, Bj, Cj);
  SB_AT(W.Aj, 2 * t) = Aj[0]; SB_AT(W.Aj, 2 * t + 1) = Aj[1];
  for (int k = 0; k < 6; ++k) SB_AT(W.Bj, 6 * t + k) = Bj[k];
  for (int k = 0; k < 3; ++k) SB_AT(W.Cj, 3 * t + k) = Cj[k];
  SB_AT(W.pref, 2 * t) = pref[(size_t)b * 2 * T + t];
  SB_AT(W.pref, 2 * t + 1) = pref[(size_t)b * 2 * T + T + t];
  const double c = cos(st[2]), s = sin(st[2]);
  double skk = 0, sgk = 0;
  const size_t NT = (size_t)N * T;
  const float* cf = coef + (size_t)b * 5 * NT;
  for (int o = 0; o < N; ++o) {
    const size_t i = (size_t)o * T + t;
    const float axf = cf[i], ayf = cf[NT + i];
    SB_AT(W.hx, i) = axf; SB_AT(W.hy, i) = ayf; SB_AT(W.hc, i) = cf[2 * NT + i];
    const double ax = axf, ay = ayf;
    const double k0 = -ax * s + ay * c, k1 = -ax * c - ay * s;          // a R'
    const double g0 = (double)cf[3 * NT + i] + ax * c + ay * s;         // mu'G + xi + a R
    const double g1 = (double)cf[4 * NT + i] - ax * s + ay * c;
    skk += k0 * k0 + k1 * k1;
    sgk += g0 * k0 + g1 * k1;
  }
  SB_AT(W.Skk, t) = skk; SB_AT(W.Sgk, t) = sgk;
  SB_AT(W.u[0], 2 * t) = ut[0]; SB_AT(W.u[0], 2 * t + 1) = ut[1];
  SB_AT(W.dd[0], t) = (double)dis[(size_t)b * T + t];
}


// Rebuild the ascending list of active instances (after the set-up and after every convergence test): the
// per-iteration kernels index through it, so their warps stay full while instances drop out.
#if defined(RDA_SB_EMULATE)
__global__ void ksb_compact(SuBatch W) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int n = 0;
  for (int b = 0; b < W.nb; ++b) if (W.active[b]) W.list[n++] = b;
  *W.n_active = n;
}
#else
__global__ void __launch_bounds__(1024) ksb_compact(SuBatch W) {
  __shared__ int wcnt[32];
  __shared__ int base, total;
  const int nb = W.nb, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int start = 0; start < nb; start += 1024) {
    const int b = start + threadIdx.x;
    const bool flag = b < nb && W.active[b] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) wcnt[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
      const int v = wcnt[lane];
      int incl = v;
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
      wcnt[lane] = incl - v;
      if (lane == 31) total = incl;
    }
    __syncthreads();
    if (flag) W.list[base + wcnt[warp] + __popc(m & ((1u << lane) - 1u))] = b;
    __syncthreads();
    if (threadIdx.x == 0) base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *W.n_active = base;
}
#endif

// initial iterate: roll the linearised model out from s_0 (serial over the horizon)
__global__ void ksb_rollout(SuBatch W) {
  const int nb = W.nb, T = W.T;
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= *W.n_active) return;
  const int b = W.list[li];
  double s0 = SB_AT(W.lins, 0), s1 = SB_AT(W.lins, 1), s2 = SB_AT(W.lins, 2);
  SB_AT(W.s[0], 0) = s0; SB_AT(W.s[0], 1) = s1; SB_AT(W.s[0], 2) = s2;
  SB_AT(W.s[1], 0) = s0; SB_AT(W.s[1], 1) = s1; SB_AT(W.s[1], 2) = s2;
  for (int t = 0; t < T; ++t) {
    const double u0 = SB_AT(W.u[0], 2 * t), u1 = SB_AT(W.u[0], 2 * t + 1);
    const double n0 = s0 + SB_AT(W.Aj, 2 * t) * s2 + SB_AT(W.Bj, 6 * t) * u0 + SB_AT(W.Bj, 6 * t + 1) * u1 + SB_AT(W.Cj, 3 * t);
    const double n1 = s1 + SB_AT(W.Aj, 2 * t + 1) * s2 + SB_AT(W.Bj, 6 * t + 2) * u0 + SB_AT(W.Bj, 6 * t + 3) * u1 + SB_AT(W.Cj, 3 * t + 1);
    const double n2 = s2 + SB_AT(W.Bj, 6 * t + 4) * u0 + SB_AT(W.Bj, 6 * t + 5) * u1 + SB_AT(W.Cj, 3 * t + 2);
    s0 = n0; s1 = n1; s2 = n2;
    SB_AT(W.s[0], 3 * t + 3) = s0; SB_AT(W.s[0], 3 * t + 4) = s1; SB_AT(W.s[0], 3 * t + 5) = s2;
  }
}

// ---- update of iteration it-1 fused with the predictor assembly of iteration it ------------------
__global__ void __launch_bounds__(128) ksb_assemble(SuBatch W, SuParams P, int it) {
  const int nb = W.nb, T = W.T, N = W.N;
  const int li = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (li >= *W.n_active) return;
  const int b = W.list[li];
  const float* __restrict__ hx_ = W.hx; const float* __restrict__ hy_ = W.hy; const float* __restrict__ hc_ = W.hc;
  double* __restrict__ hs_ = W.hs; double* __restrict__ hnu_ = W.hnu;
  const bool accm = P.accelerated != 0;
  const double ro1 = P.ro1, ro2 = P.ro2, iro1 = 1.0 / ro1, reg = 1e-9;
  const int cur = it & 1, nxt = cur ^ 1;
  // old iterate of this stage (and the previous stage's controls for the rate rows)
  double ut[2] = {SB_AT(W.u[cur], 2 * t), SB_AT(W.u[cur], 2 * t + 1)};
  double up[2] = {0, 0};
  if (t >= 1) { up[0] = SB_AT(W.u[cur], 2 * t - 2); up[1] = SB_AT(W.u[cur], 2 * t - 1); }
  double dt_ = SB_AT(W.dd[cur], t);
  double sn[3] = {SB_AT(W.s[cur], 3 * t + 3), SB_AT(W.s[cur], 3 * t + 4), SB_AT(W.s[cur], 3 * t + 5)};
  const double pr0 = SB_AT(W.pref, 2 * t), pr1 = SB_AT(W.pref, 2 * t + 1);
  double a = 0, sigma_mu = 0;
  double dvt[3] = {0, 0, 0}, dup[2] = {0, 0}, dvat[3] = {0, 0, 0}, dupa[2] = {0, 0};
  double dzn[3] = {0, 0, 0}, dzan[2] = {0, 0};
  const double mu0 = P.mu0 > 0 ? (double)P.mu0 : 1.0;
  if (it > 0) {
    a = W.alpha[b];
    sigma_mu = W.sigma_mu[b];
    for (int k = 0; k < 3; ++k) { dvt[k] = SB_AT(W.dv, 3 * t + k); dvat[k] = SB_AT(W.dva, 3 * t + k); }
    for (int k = 0; k < 2; ++k) { dup[k] = SB_AT(W.dz, 5 * t + 3 + k); dupa[k] = SB_AT(W.dza, 5 * t + 3 + k); }
    for (int k = 0; k < 3; ++k) dzn[k] = SB_AT(W.dz, 5 * t + 5 + k);
    for (int k = 0; k < 2; ++k) dzan[k] = SB_AT(W.dza, 5 * t + 5 + k);
  }
  // new iterate of this stage
  double utn[2] = {ut[0] + a * dvt[0], ut[1] + a * dvt[1]};
  double upn[2] = {up[0] + a * dup[0], up[1] + a * dup[1]};
  double dtn = N > 0 ? dt_ + a * dvt[2] : dt_;
  double snn[3] = {sn[0] + a * dzn[0], sn[1] + a * dzn[1], sn[2] + a * dzn[2]};
  SB_AT(W.p_step, t) = fmax(fabs(dvt[0]), fmax(fabs(dvt[1]), fabs(dvt[2])));
  SB_AT(W.u[nxt], 2 * t) = utn[0]; SB_AT(W.u[nxt], 2 * t + 1) = utn[1];
  SB_AT(W.dd[nxt], t) = dtn;
  SB_AT(W.s[nxt], 3 * t + 3) = snn[0]; SB_AT(W.s[nxt], 3 * t + 4) = snn[1]; SB_AT(W.s[nxt], 3 * t + 5) = snn[2];
  // ---- stage gradient of the smooth part at the new iterate ----
  const double tw = 2 * (double)P.ws;
  double gw[8];
  gw[0] = tw * (snn[0] - (double)SB_AT(W.ref, 3 * t + 3));
  gw[1] = tw * (snn[1] - (double)SB_AT(W.ref, 3 * t + 4));
  gw[2] = (P.dynamics == RDA_DYN_OMNI ? 0.0 : tw * (snn[2] - (double)SB_AT(W.ref, 3 * t + 5)))
          + ro2 * (SB_AT(W.Skk, t) * (snn[2] - (double)SB_AT(W.lins, 3 * t + 2)) + SB_AT(W.Sgk, t));
  gw[3] = 2 * (double)P.wu * (utn[0] - W.vref[b]) + reg * utn[0];
  gw[4] = reg * utn[1];
  gw[5] = N > 0 ? -(double)P.slack_gain + reg * dtn : 0.0;
  gw[6] = 0; gw[7] = 0;
  double wb[5] = {0, 0, 0, 0, 0};
  double acc_mu = 0, acc_r = 0;
  // ---- box / rate rows: update (old iterate), then predictor terms (new iterate) ----
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    const SbRow ro = sb_row(P, t, c, ut, up, dt_);
    double sv, nu;
    if (it == 0) {
      sv = ro.live ? fmax(ro.g, 1e-2) : 1.0;
      nu = ro.live ? mu0 / sv : 0.0;
    } else {
      sv = SB_AT(W.bs, 10 * t + c); nu = SB_AT(W.bnu, 10 * t + c);
      if (ro.live) {
        const double res = ro.g - sv;
        const double dir = sb_row_dir(ro, dvt, dup), dira = sb_row_dir(ro, dvat, dupa);
        const double ds = dir + res, dsa = dira + res;
        const double isv = rcp_(sv), om = nu * isv;
        const double dna = -nu - om * dsa;
        const double dn = (sigma_mu - dsa * dna) * isv - nu - om * ds;
        sv += a * ds; nu += a * dn;
      }
    }
    SB_AT(W.bs, 10 * t + c) = sv; SB_AT(W.bnu, 10 * t + c) = nu;
    if (!ro.live) continue;
    const SbRow r = sb_row(P, t, c, utn, upn, dtn);
    const double res = r.g - sv;
    const double isv = rcp_(sv), om = nu * isv;
    const double term = -om * res;
    acc_mu += sv * nu;
    acc_r = fmax(acc_r, fabs(res));
    gw[r.comp] -= r.sgn * term;
    if (r.rate) gw[r.comp + 3] += r.sgn * term;
    wb[(r.rate ? 3 : 0) + (r.comp - 3)] += om;
  }
  // ---- hinges ----
  double m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0;
  const double dxo = sn[0] - pr0, dyo = sn[1] - pr1;        // old position offset
  const double dxn = snn[0] - pr0, dyn = snn[1] - pr1;
  double g0 = gw[0], g1 = gw[1], g5 = gw[5];
#pragma unroll 4
  for (int o = 0; o < N; ++o) {
    const size_t i = (size_t)o * T + t;
    const double ax = SB_AT(hx_, i), ay = SB_AT(hy_, i), hc = SB_AT(hc_, i);
    double tk, om;
    if (accm) {
      double sv, nu;
      if (it == 0) {
        const double l = ax * dxn + ay * dyn + hc - dtn;
        sv = (l + sqrt(l * l + 4 * mu0 / ro1)) / 2;
        nu = mu0 / sv;
      } else {
        sv = SB_AT(hs_, i); nu = SB_AT(hnu_, i);
        const double l = ax * dxo + ay * dyo + hc - dt_;
        const double nr = nu * iro1;
        const double res = l + nr - sv;
        const double iden = rcp_(sv + nr);
        const double om_ = nu * iden;
        const double dir = ax * dzn[0] + ay * dzn[1] - dvt[2];
        const double dira = ax * dzan[0] + ay * dzan[1] - dvat[2];
        const double dna = -om_ * (sv + res + dira);
        const double dsa = dira + dna * iro1 + res;
        const double cc = sv * nu - sigma_mu + dsa * dna;
        const double dn = -(cc + nu * res + nu * dir) * iden;
        const double ds = dir + dn * iro1 + res;
        sv += a * ds; nu += a * dn;
      }
      SB_AT(hs_, i) = sv; SB_AT(hnu_, i) = nu;
      const double l = ax * dxn + ay * dyn + hc - dtn;
      const double nr = nu * iro1;
      const double res = l + nr - sv;
      const double iden = rcp_(sv + nr);
      om = nu * iden;
      tk = om * (nr - res);
      acc_mu += sv * nu;
      acc_r = fmax(acc_r, fabs(res));
    } else {
      const double l = ax * dxn + ay * dyn + hc - dtn;
      om = ro1;
      tk = -ro1 * l;
    }
    g0 -= ax * tk; g1 -= ay * tk; g5 += tk;
    m0 += om * ax * ax; m1 += om * ax * ay; m2 -= om * ax;
    m3 += om * ay * ay; m4 -= om * ay; m5 += om;
  }
  // eliminate d_t (Schur complement on Q_dd)
  const double iq = N > 0 ? rcp_(reg + wb[2] + m5) : 1.0;
  const double e0 = m2 * iq, e1 = m4 * iq;
  SB_AT(W.Wm, 3 * t) = m0 - m2 * e0; SB_AT(W.Wm, 3 * t + 1) = m1 - m2 * e1; SB_AT(W.Wm, 3 * t + 2) = m3 - m4 * e1;
  SB_AT(W.Ed, 3 * t) = e0; SB_AT(W.Ed, 3 * t + 1) = e1; SB_AT(W.Ed, 3 * t + 2) = iq;
  SB_AT(W.g5q, t) = g5 * iq;
  gw[0] = g0 - e0 * g5; gw[1] = g1 - e1 * g5; gw[5] = g5;
  for (int k = 0; k < 8; ++k) SB_AT(W.gw, 8 * t + k) = gw[k];
  for (int k = 0; k < 5; ++k) SB_AT(W.wb, 5 * t + k) = wb[k];
  SB_AT(W.p_mu, t) = acc_mu;
  SB_AT(W.p_r, t) = acc_r;
}

// ---- Riccati recursion, one thread per instance ---------------------------------------------------
// FACTOR: convergence test + factorisation + forward sweep into (dza, dva); else solve + forward into (dz, dv)
struct SbOut { float *cur_s, *cur_u, *dis; int *status, *iters, *counters; };

template <bool FACTOR>
__global__ void __launch_bounds__(64) ksb_riccati(SuBatch W, SuParams P, int it, double Mrows, SbOut out) {
  const int nb = W.nb, T = W.T;
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= *W.n_active) return;
  const int b = W.list[li];
  const double reg = 1e-9;
  if (FACTOR) {
    // per-instance reductions over the horizon in a fixed order (deterministic)
    double smu = 0, rmx = 0, stp = 0;
    for (int t = 0; t < T; ++t) {
      smu += SB_AT(W.p_mu, t);
      rmx = fmax(rmx, SB_AT(W.p_r, t));
      stp = fmax(stp, SB_AT(W.p_step, t));
    }
    const double mu = smu / Mrows;
    const double last_step = it > 0 ? W.alpha[b] * stp : 1e30;
    W.mu[b] = mu;
    int fin = -1;
    if (!isfinite(mu)) fin = 2;
    else if (mu < 1e-9 && rmx < 1e-9 && (last_step < 1e-6 || mu < 1e-13)) fin = 0;
    else if (it >= P.max_iter) fin = 1;
    if (fin >= 0) {
      // the current iterate is in buffer (it + 1) & 1; accept OPTIMAL / OPTIMAL_INACCURATE (iteration cap),
      // else keep the previous nominal (rda_solver.py:696-700)
      const int cb = (it + 1) & 1;
      bool ok = fin != 2;
      if (ok) {
        for (int i = 0; i < 3 * (T + 1) && ok; ++i) ok = isfinite((float)SB_AT(W.s[cb], i));
        for (int i = 0; i < 2 * T && ok; ++i) ok = isfinite((float)SB_AT(W.u[cb], i));
      }
      if (ok) {
        float* ws = out.cur_s + (size_t)b * 3 * (T + 1);
        float* wu = out.cur_u + (size_t)b * 2 * T;
        for (int t = 0; t <= T; ++t)
          for (int r = 0; r < 3; ++r) ws[r * (T + 1) + t] = (float)SB_AT(W.s[cb], 3 * t + r);
        for (int t = 0; t < T; ++t) {
          wu[t] = (float)SB_AT(W.u[cb], 2 * t); wu[T + t] = (float)SB_AT(W.u[cb], 2 * t + 1);
          out.dis[(size_t)b * T + t] = (float)SB_AT(W.dd[cb], t);
        }
      }
      int flag = 0;
      if (fin == 1) flag |= RDA_ST_SU_NOT_CONVERGED;
      if (!ok) flag |= RDA_ST_SU_NONFINITE;
      if (flag) out.status[b] |= flag;
      out.iters[b] += 1;
      atomicAdd(&out.counters[3], it);
      atomicAdd(&out.counters[4], 1);
      W.active[b] = 0; W.st[b] = fin; W.its[b] = it;        // ksb_compact drops it from the list
      return;
    }
  }
  const double tw = 2 * (double)P.ws, tw3 = (P.dynamics == RDA_DYN_OMNI ? 0.0 : tw);
  double Pm[5][5], pv[5];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    pv[a] = 0;
#pragma unroll
    for (int c = 0; c < 5; ++c) Pm[a][c] = 0;
  }
  // operands of one stage, prefetched one stage ahead
  double nA[2], nB[6], nG[7], nM[3], nW[5] = {0, 0, 0, 0, 0}, nK[10], nS = 0;
  auto load_stage = [&](int t) {
    nA[0] = SB_AT(W.Aj, 2 * t); nA[1] = SB_AT(W.Aj, 2 * t + 1);
#pragma unroll
    for (int k = 0; k < 6; ++k) nB[k] = SB_AT(W.Bj, 6 * t + k);
#pragma unroll
    for (int k = 0; k < 5; ++k) nG[k] = SB_AT(W.gw, 8 * t + k);
    nG[5] = SB_AT(W.gw, 8 * t + 6); nG[6] = SB_AT(W.gw, 8 * t + 7);
    if (FACTOR) {
#pragma unroll
      for (int k = 0; k < 3; ++k) nM[k] = SB_AT(W.Wm, 3 * t + k);
#pragma unroll
      for (int k = 0; k < 5; ++k) nW[k] = SB_AT(W.wb, 5 * t + k);
      nS = SB_AT(W.Skk, t);
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) nM[k] = SB_AT(W.Lc, 3 * t + k);
#pragma unroll
      for (int k = 0; k < 10; ++k) nK[k] = SB_AT(W.K, 10 * t + k);
    }
  };
  load_stage(T - 1);
  for (int t = T - 1; t >= 0; --t) {
    const double a02 = nA[0], a12 = nA[1];
    const double b00 = nB[0], b01 = nB[1], b10 = nB[2], b11 = nB[3], b20 = nB[4], b21 = nB[5];
    const double G0 = nG[0], G1 = nG[1], G2 = nG[2], G3 = nG[3], G4 = nG[4], G6 = nG[5], G7 = nG[6];
    const double M0 = nM[0], M1 = nM[1], M2 = nM[2];
    const double w0 = nW[0], w1 = nW[1], wr0 = nW[3], wr1 = nW[4], skk = nS;
    double Kt[2][5];
    if (!FACTOR) {
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int c = 0; c < 5; ++c) Kt[k][c] = nK[5 * k + c];
    }
    if (t > 0) load_stage(t - 1);
    const double q0 = G0 + pv[0], q1 = G1 + pv[1], q2 = G2 + pv[2], q3 = G3 + pv[3], q4 = G4 + pv[4];
    const double gz0 = q0, gz1 = q1, gz2 = a02 * q0 + a12 * q1 + q2, gz3 = G6, gz4 = G7;
    const double gv0 = b00 * q0 + b10 * q1 + b20 * q2 + q3;
    const double gv1 = b01 * q0 + b11 * q1 + b21 * q2 + q4;
    double i00, L10, i11;
    if (FACTOR) {
      double Q[5][5];
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int c = 0; c < 5; ++c) Q[a][c] = Pm[a][c];
      Q[0][0] += tw + M0; Q[0][1] += M1; Q[1][0] += M1; Q[1][1] += tw + M2;
      Q[2][2] += tw3 + (double)P.ro2 * skk;
      Q[3][3] += 2 * (double)P.wu + reg + w0 + wr0;
      Q[4][4] += reg + w1 + wr1;
      double t2[5], t5[5], t6[5];
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        t2[r] = a02 * Q[r][0] + a12 * Q[r][1] + Q[r][2];
        t5[r] = b00 * Q[r][0] + b10 * Q[r][1] + b20 * Q[r][2] + Q[r][3];
        t6[r] = b01 * Q[r][0] + b11 * Q[r][1] + b21 * Q[r][2] + Q[r][4];
      }
#define SB_J2(x) (a02 * (x)[0] + a12 * (x)[1] + (x)[2])
#define SB_J5(x) (b00 * (x)[0] + b10 * (x)[1] + b20 * (x)[2] + (x)[3])
#define SB_J6(x) (b01 * (x)[0] + b11 * (x)[1] + b21 * (x)[2] + (x)[4])
      double Hzz[5][5];
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int c = 0; c < 5; ++c) Hzz[a][c] = 0;
      Hzz[0][0] = Q[0][0]; Hzz[0][1] = Q[0][1]; Hzz[1][1] = Q[1][1];
      Hzz[0][2] = t2[0]; Hzz[1][2] = t2[1]; Hzz[2][2] = SB_J2(t2);
      Hzz[3][3] = wr0; Hzz[4][4] = wr1;
      double Hvz[2][5];
      Hvz[0][0] = t5[0]; Hvz[0][1] = t5[1]; Hvz[0][2] = SB_J2(t5); Hvz[0][3] = -wr0; Hvz[0][4] = 0;
      Hvz[1][0] = t6[0]; Hvz[1][1] = t6[1]; Hvz[1][2] = SB_J2(t6); Hvz[1][3] = 0; Hvz[1][4] = -wr1;
      const double h00 = SB_J5(t5), h10 = SB_J6(t5), h11 = SB_J6(t6);
#undef SB_J2
#undef SB_J5
#undef SB_J6
      i00 = rcp_(h00);
      L10 = h10 * i00;
      i11 = rcp_(h11 - L10 * h10);
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const double y0 = -Hvz[0][c];
        const double y1 = -Hvz[1][c] - L10 * y0;
        const double x1 = y1 * i11;
        const double x0 = y0 * i00 - L10 * x1;
        Kt[0][c] = x0; Kt[1][c] = x1;
      }
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int c = a; c < 5; ++c) {
          const double v = Hzz[a][c] + Hvz[0][a] * Kt[0][c] + Hvz[1][a] * Kt[1][c];
          Pm[a][c] = v; Pm[c][a] = v;
        }
      SB_AT(W.Lc, 3 * t) = i00; SB_AT(W.Lc, 3 * t + 1) = L10; SB_AT(W.Lc, 3 * t + 2) = i11;
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int c = 0; c < 5; ++c) SB_AT(W.K, 10 * t + 5 * k + c) = Kt[k][c];
    } else {
      i00 = M0; L10 = M1; i11 = M2;
    }
    {
      const double y0 = -gv0;
      const double y1 = -gv1 - L10 * y0;
      const double x1 = y1 * i11;
      const double x0 = y0 * i00 - L10 * x1;
      SB_AT(W.kf, 2 * t) = x0; SB_AT(W.kf, 2 * t + 1) = x1;
      pv[0] = gz0 + Kt[0][0] * gv0 + Kt[1][0] * gv1;
      pv[1] = gz1 + Kt[0][1] * gv0 + Kt[1][1] * gv1;
      pv[2] = gz2 + Kt[0][2] * gv0 + Kt[1][2] * gv1;
      pv[3] = gz3 + Kt[0][3] * gv0 + Kt[1][3] * gv1;
      pv[4] = gz4 + Kt[0][4] * gv0 + Kt[1][4] * gv1;
    }
  }
  // forward sweep (+ recovery of the eliminated d step)
  double* dzp = FACTOR ? W.dza : W.dz;
  double* dvp = FACTOR ? W.dva : W.dv;
  double z[5] = {0, 0, 0, 0, 0};
  double fK[10], fk[2], fA[2], fB[6], fE[2], fq;
  auto load_fwd = [&](int t) {
#pragma unroll
    for (int k = 0; k < 10; ++k) fK[k] = SB_AT(W.K, 10 * t + k);
    fk[0] = SB_AT(W.kf, 2 * t); fk[1] = SB_AT(W.kf, 2 * t + 1);
    fA[0] = SB_AT(W.Aj, 2 * t); fA[1] = SB_AT(W.Aj, 2 * t + 1);
#pragma unroll
    for (int k = 0; k < 6; ++k) fB[k] = SB_AT(W.Bj, 6 * t + k);
    fE[0] = SB_AT(W.Ed, 3 * t); fE[1] = SB_AT(W.Ed, 3 * t + 1); fq = SB_AT(W.g5q, t);
  };
  load_fwd(0);
  for (int t = 0; t < T; ++t) {
    double v[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      double sacc = fk[k];
#pragma unroll
      for (int c = 0; c < 5; ++c) sacc += fK[5 * k + c] * z[c];
      v[k] = sacc;
    }
    const double n0 = z[0] + fA[0] * z[2] + fB[0] * v[0] + fB[1] * v[1];
    const double n1 = z[1] + fA[1] * z[2] + fB[2] * v[0] + fB[3] * v[1];
    const double n2 = z[2] + fB[4] * v[0] + fB[5] * v[1];
    const double e0 = fE[0], e1 = fE[1], q5 = fq;
#pragma unroll
    for (int a = 0; a < 5; ++a) SB_AT(dzp, 5 * t + a) = z[a];
    if (t + 1 < T) load_fwd(t + 1);
    SB_AT(dvp, 3 * t) = v[0]; SB_AT(dvp, 3 * t + 1) = v[1];
    SB_AT(dvp, 3 * t + 2) = P.N > 0 ? -(q5 + e0 * n0 + e1 * n1) : 0.0;
    z[0] = n0; z[1] = n1; z[2] = n2; z[3] = v[0]; z[4] = v[1];
  }
#pragma unroll
  for (int a = 0; a < 5; ++a) SB_AT(dzp, 5 * T + a) = z[a];
}

// ---- step length and predicted complementarity of the affine direction ----------------------------
// PHASE 0: affine direction (dza, dva): max step ratio and the three sums of the complementarity
// polynomial.  PHASE 1: combined direction (dz, dv): max step ratio only.
template <int PHASE>
__global__ void __launch_bounds__(128) ksb_steplen(SuBatch W, SuParams P, int it) {
  const int nb = W.nb, T = W.T, N = W.N;
  const int li = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (li >= *W.n_active) return;
  const int b = W.list[li];
  const float* __restrict__ hx_ = W.hx; const float* __restrict__ hy_ = W.hy; const float* __restrict__ hc_ = W.hc;
  const double* __restrict__ hs_ = W.hs; const double* __restrict__ hnu_ = W.hnu;
  const bool accm = P.accelerated != 0;
  const double iro1 = 1.0 / (double)P.ro1;
  const int cb = (it + 1) & 1;
  const double ut[2] = {SB_AT(W.u[cb], 2 * t), SB_AT(W.u[cb], 2 * t + 1)};
  double up[2] = {0, 0};
  if (t >= 1) { up[0] = SB_AT(W.u[cb], 2 * t - 2); up[1] = SB_AT(W.u[cb], 2 * t - 1); }
  const double dt_ = SB_AT(W.dd[cb], t);
  const double dx = SB_AT(W.s[cb], 3 * t + 3) - (double)SB_AT(W.pref, 2 * t);
  const double dy = SB_AT(W.s[cb], 3 * t + 4) - (double)SB_AT(W.pref, 2 * t + 1);
  double dvat[3], dupa[2], dzan[2];
  for (int k = 0; k < 3; ++k) dvat[k] = SB_AT(W.dva, 3 * t + k);
  for (int k = 0; k < 2; ++k) { dupa[k] = SB_AT(W.dza, 5 * t + 3 + k); dzan[k] = SB_AT(W.dza, 5 * t + 5 + k); }
  double dvt[3] = {0, 0, 0}, dup[2] = {0, 0}, dzn[2] = {0, 0};
  double sigma_mu = 0;
  if (PHASE == 1) {
    for (int k = 0; k < 3; ++k) dvt[k] = SB_AT(W.dv, 3 * t + k);
    for (int k = 0; k < 2; ++k) { dup[k] = SB_AT(W.dz, 5 * t + 3 + k); dzn[k] = SB_AT(W.dz, 5 * t + 5 + k); }
    sigma_mu = W.sigma_mu[b];
  }
  double rmaxr = 0, s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    const SbRow r = sb_row(P, t, c, ut, up, dt_);
    if (!r.live) continue;
    const double sv = SB_AT(W.bs, 10 * t + c), nu = SB_AT(W.bnu, 10 * t + c);
    const double res = r.g - sv;
    const double ip = rcp_(sv * nu);
    const double isv = nu * ip, inu = sv * ip, om = nu * isv;
    const double dira = sb_row_dir(r, dvat, dupa);
    double ds, dn;
    if (PHASE == 0) {
      ds = dira + res;
      dn = -nu - om * ds;
    } else {
      const double dir = sb_row_dir(r, dvt, dup);
      ds = dir + res;
      const double dsa = dira + res;
      const double dna = -nu - om * dsa;
      dn = (sigma_mu - dsa * dna) * isv - nu - om * ds;
    }
    rmaxr = fmax(rmaxr, fmax(-ds * isv, -dn * inu));
    if (PHASE == 0) { s0 += sv * nu; s1 += sv * dn + nu * ds; s2 += ds * dn; }
  }
  if (accm) {
#pragma unroll 4
    for (int o = 0; o < N; ++o) {
      const size_t i = (size_t)o * T + t;
      const double ax = SB_AT(hx_, i), ay = SB_AT(hy_, i);
      const double l = ax * dx + ay * dy + (double)SB_AT(hc_, i) - dt_;
      const double sv = SB_AT(hs_, i), nu = SB_AT(hnu_, i);
      const double nr = nu * iro1;
      const double res = l + nr - sv;
      const double iden = rcp_(sv + nr);
      const double om = nu * iden;
      const double dira = ax * dzan[0] + ay * dzan[1] - dvat[2];
      double dir = dira, cc = sv * nu;
      if (PHASE == 1) {
        dir = ax * dzn[0] + ay * dzn[1] - dvt[2];
        const double dna = -om * (sv + res + dira);
        const double dsa = dira + dna * iro1 + res;
        cc = sv * nu - sigma_mu + dsa * dna;
      }
      const double dn = -(cc + nu * res + nu * dir) * iden;
      const double ds = dir + dn * iro1 + res;
      const double ip = rcp_(sv * nu);
      rmaxr = fmax(rmaxr, fmax(-ds * nu * ip, -dn * sv * ip));
      if (PHASE == 0) { s0 += sv * nu; s1 += sv * dn + nu * ds; s2 += ds * dn; }
    }
  }
  SB_AT(W.p_rmax, t) = rmaxr;
  if (PHASE == 0) { SB_AT(W.p_s0, t) = s0; SB_AT(W.p_s1, t) = s1; SB_AT(W.p_s2, t) = s2; }
}

// per-instance reductions of ksb_steplen (fixed order): PHASE 0 -> sigma * mu, PHASE 1 -> step length
template <int PHASE>
__global__ void __launch_bounds__(128) ksb_reduce(SuBatch W, double Mrows) {
  const int nb = W.nb, T = W.T;
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= *W.n_active) return;
  const int b = W.list[li];
  double rm = 0, s0 = 0, s1 = 0, s2 = 0;
  for (int t = 0; t < T; ++t) {
    rm = fmax(rm, SB_AT(W.p_rmax, t));
    if (PHASE == 0) { s0 += SB_AT(W.p_s0, t); s1 += SB_AT(W.p_s1, t); s2 += SB_AT(W.p_s2, t); }
  }
  const double amax = rm > 1e-30 ? 1.0 / rm : 1e30;
  if (PHASE == 0) {
    const double mu = W.mu[b];
    const double a = fmin(1.0, amax);
    const double mua = (s0 + a * s1 + a * a * s2) / Mrows;
    double sg = mua / mu;
    sg = sg * sg * sg;
    W.sigma_mu[b] = fmin(sg, 1.0) * mu;
  } else {
    // fraction to the boundary as su_solve's (su_solver.cuh, RDA_SU_TAU_ADAPT)
    const double tau_b = (RDA_SU_TAU_ADAPT) > 0 ? fmax(0.995, 1.0 - (double)(RDA_SU_TAU_ADAPT) * W.mu[b]) : 0.995;
    W.alpha[b] = fmin(1.0, tau_b * amax);
  }
}

// ---- corrector right-hand side (needs sigma * mu of the instance) ----------------------------------
__global__ void __launch_bounds__(128) ksb_corrector(SuBatch W, SuParams P, int it) {
  const int nb = W.nb, T = W.T, N = W.N;
  const int li = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (li >= *W.n_active) return;
  const int b = W.list[li];
  const float* __restrict__ hx_ = W.hx; const float* __restrict__ hy_ = W.hy; const float* __restrict__ hc_ = W.hc;
  const double* __restrict__ hs_ = W.hs; const double* __restrict__ hnu_ = W.hnu;
  const bool accm = P.accelerated != 0;
  const double ro2 = P.ro2, iro1 = 1.0 / (double)P.ro1, reg = 1e-9;
  const int cb = (it + 1) & 1;
  const double sigma_mu = W.sigma_mu[b];        // ksb_reduce<0>
  const double ut[2] = {SB_AT(W.u[cb], 2 * t), SB_AT(W.u[cb], 2 * t + 1)};
  double up[2] = {0, 0};
  if (t >= 1) { up[0] = SB_AT(W.u[cb], 2 * t - 2); up[1] = SB_AT(W.u[cb], 2 * t - 1); }
  const double dt_ = SB_AT(W.dd[cb], t);
  const double sn[3] = {SB_AT(W.s[cb], 3 * t + 3), SB_AT(W.s[cb], 3 * t + 4), SB_AT(W.s[cb], 3 * t + 5)};
  double dvat[3], dupa[2], dzan[2];
  for (int k = 0; k < 3; ++k) dvat[k] = SB_AT(W.dva, 3 * t + k);
  for (int k = 0; k < 2; ++k) { dupa[k] = SB_AT(W.dza, 5 * t + 3 + k); dzan[k] = SB_AT(W.dza, 5 * t + 5 + k); }
  const double tw = 2 * (double)P.ws;
  double gw[8];
  gw[0] = tw * (sn[0] - (double)SB_AT(W.ref, 3 * t + 3));
  gw[1] = tw * (sn[1] - (double)SB_AT(W.ref, 3 * t + 4));
  gw[2] = (P.dynamics == RDA_DYN_OMNI ? 0.0 : tw * (sn[2] - (double)SB_AT(W.ref, 3 * t + 5)))
          + ro2 * (SB_AT(W.Skk, t) * (sn[2] - (double)SB_AT(W.lins, 3 * t + 2)) + SB_AT(W.Sgk, t));
  gw[3] = 2 * (double)P.wu * (ut[0] - W.vref[b]) + reg * ut[0];
  gw[4] = reg * ut[1];
  gw[5] = N > 0 ? -(double)P.slack_gain + reg * dt_ : 0.0;
  gw[6] = 0; gw[7] = 0;
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    const SbRow r = sb_row(P, t, c, ut, up, dt_);
    if (!r.live) continue;
    const double sv = SB_AT(W.bs, 10 * t + c), nu = SB_AT(W.bnu, 10 * t + c);
    const double res = r.g - sv;
    const double isv = rcp_(sv), om = nu * isv;
    const double dsa = sb_row_dir(r, dvat, dupa) + res;
    const double dna = -nu - om * dsa;
    const double term = (sigma_mu - dsa * dna) * isv - om * res;
    gw[r.comp] -= r.sgn * term;
    if (r.rate) gw[r.comp + 3] += r.sgn * term;
  }
  const double dx = sn[0] - (double)SB_AT(W.pref, 2 * t), dy = sn[1] - (double)SB_AT(W.pref, 2 * t + 1);
  double g0 = gw[0], g1 = gw[1], g5 = gw[5];
#pragma unroll 4
  for (int o = 0; o < N; ++o) {
    const size_t i = (size_t)o * T + t;
    const double ax = SB_AT(hx_, i), ay = SB_AT(hy_, i);
    const double l = ax * dx + ay * dy + (double)SB_AT(hc_, i) - dt_;
    double tk;
    if (accm) {
      const double sv = SB_AT(hs_, i), nu = SB_AT(hnu_, i);
      const double nr = nu * iro1;
      const double res = l + nr - sv;
      const double iden = rcp_(sv + nr);
      const double om = nu * iden;
      const double dir = ax * dzan[0] + ay * dzan[1] - dvat[2];
      const double dna = -om * (sv + res + dir);
      const double dsa = dir + dna * iro1 + res;
      const double cc = sv * nu - sigma_mu + dsa * dna;
      tk = nu - (cc + nu * res) * iden;
    } else {
      tk = -(double)P.ro1 * l;
    }
    g0 -= ax * tk; g1 -= ay * tk; g5 += tk;
  }
  const double e0 = SB_AT(W.Ed, 3 * t), e1 = SB_AT(W.Ed, 3 * t + 1), iq = SB_AT(W.Ed, 3 * t + 2);
  SB_AT(W.g5q, t) = g5 * iq;
  gw[0] = g0 - e0 * g5; gw[1] = g1 - e1 * g5; gw[5] = g5;
  for (int k = 0; k < 8; ++k) SB_AT(W.gw, 8 * t + k) = gw[k];
}

#undef SB_AT
#endif  // __CUDACC__ || RDA_SB_EMULATE

}  // namespace rda
