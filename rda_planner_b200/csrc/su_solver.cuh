// su_solver.cuh — the state/control "su" QP of one planning instance.
//
// Replaces rda_solver.py: su_prob_solve :692-700 (cvxpy -> ECOS) for the problem defined by
// construct_su_prob :216-231, nav_cost_cons :313-328, update_su_cost_cons :330-387,
// Im_su :831-851, Hm_su :853-872, dynamics_constraint :911-928, bound_su_constraints :930-938,
// bound_dis_constraints :940-947, C0_cost :1011-1029, C1_cost :1031-1032, and the
// re-linearisation assign_state_parameter :436-460 / linear_*_model :949-994.
//
// Method: Mehrotra predictor-corrector primal-dual interior point method.
//  * decision variables (u_t, d_t); states follow the linearised dynamics exactly, so every Newton
//    step is an equality-constrained LQ problem solved by a Riccati recursion over the horizon
//    with state (s_t, u_{t-1}) in R^5 and control (u_t, d_t) in R^3 (the banded KKT system);
//  * inequality constraints: |u| <= max_speed, |u_t - u_{t-1}| <= max_acce*dt, min_sd <= d <= max_sd;
//  * every hinge term ro1/2 neg(Im)^2 is the exact partial minimum over a slack w of
//    ro1/2 w^2 s.t. Im + w >= 0; w is eliminated analytically (w = nu/ro1), leaving one
//    (slack, multiplier) pair per hinge and a rank-one term in the stage Hessian.
// The code is written against a "cooperative group" context Ctx (lane(), nlanes(), sync(),
// min/sum reductions): one warp per instance on the GPU (lanes = horizon stages), one lane on the
// host (tests/host_shim).
#pragma once
#include "rda_hd.h"

namespace rda {

struct SuParams {
  int T, N, dynamics, accelerated;
  float dt, L, umax[2], ab[2], ws, wu;
  float slack_gain, dmin, dmax, ro1, ro2;
  int max_iter;
  float mu0;      // initial complementarity of the interior point iteration
};

// Per-instance workspace (shared memory on the GPU).  All arrays indexed by stage t (0..T-1)
// unless noted; hinge arrays indexed [o*T + t].
template <typename Real>
struct SuWork {
  Real *s, *u, *d;            // iterate: 3(T+1), 2T, T
  Real *ref;                  // 3(T+1)
  Real *lins, *linu;          // linearisation point 3(T+1), 2T
  Real *cph, *sph;            // cos/sin of nominal heading (column t)
  Real *Aj, *Bj, *Cj;         // 2T (A02, A12), 6T, 3T
  Real *Skk, *Sgk;            // aggregated rotation-consensus terms
  Real *pref;                 // 2T positions the hinge offsets refer to
  float *hx, *hy, *hc;        // hinge rows: lam'A (2) and offset
  Real *hs, *hnu;             // hinge slack / multiplier
  Real *bs, *bnu;             // 10T box/rate slack / multiplier
  Real *Wm;                   // 6T hinge Hessian (xx, xy, xd, yy, yd, dd)
  Real *gw;                   // 8T gradient in stage coordinates
  Real *wb;                   // 5T barrier weights (u0, u1, d, rate0, rate1)
  Real *K, *Lc, *kf;          // 15T, 6T, 3T Riccati gains / Cholesky of Hvv / feed-forward
  Real *rw;                   // 128 scratch of the lane-parallel Riccati step (Q, t-columns, H blocks, P)
  Real *dz, *dv;              // 5(T+1), 3T Newton step
  Real *dza, *dva;            // affine (predictor) step
  Real vref;
};

template <typename Real>
RDA_HD size_t su_work_layout(int T, int N, SuWork<Real>* w, char* base) {
  // returns bytes used; if base != nullptr the pointers are set
  size_t off = 0;
  auto take = [&](size_t n, size_t elt) {
    off = (off + 15) & ~(size_t)15;
    char* p = base ? base + off : nullptr;
    off += n * elt;
    return p;
  };
#define RDA_TAKE(field, n, type) { char* p_ = take((size_t)(n), sizeof(type)); if (w) w->field = (type*)p_; }
  RDA_TAKE(s, 3 * (T + 1), Real) RDA_TAKE(u, 2 * T, Real) RDA_TAKE(d, T, Real)
  RDA_TAKE(ref, 3 * (T + 1), Real) RDA_TAKE(lins, 3 * (T + 1), Real) RDA_TAKE(linu, 2 * T, Real)
  RDA_TAKE(cph, T, Real) RDA_TAKE(sph, T, Real)
  RDA_TAKE(Aj, 2 * T, Real) RDA_TAKE(Bj, 6 * T, Real) RDA_TAKE(Cj, 3 * T, Real)
  RDA_TAKE(Skk, T, Real) RDA_TAKE(Sgk, T, Real) RDA_TAKE(pref, 2 * T, Real)
  RDA_TAKE(hx, N * T, float) RDA_TAKE(hy, N * T, float) RDA_TAKE(hc, N * T, float)
  RDA_TAKE(hs, N * T, Real) RDA_TAKE(hnu, N * T, Real)
  RDA_TAKE(bs, 10 * T, Real) RDA_TAKE(bnu, 10 * T, Real)
  RDA_TAKE(Wm, 6 * T, Real) RDA_TAKE(gw, 8 * T, Real) RDA_TAKE(wb, 5 * T, Real)
  RDA_TAKE(K, 15 * T, Real) RDA_TAKE(Lc, 6 * T, Real) RDA_TAKE(kf, 3 * T, Real) RDA_TAKE(rw, 128, Real)
  RDA_TAKE(dz, 5 * (T + 1), Real) RDA_TAKE(dv, 3 * T, Real)
  RDA_TAKE(dza, 5 * (T + 1), Real) RDA_TAKE(dva, 3 * T, Real)
#undef RDA_TAKE
  return (off + 15) & ~(size_t)15;
}

// Jacobians of the discrete model about (s, u): linear_ackermann_model :949-963,
// linear_diff_model :966-979, linear_omni_model :982-994.
template <typename Real>
RDA_HD void su_linearise(const SuParams& P, const Real* st, const Real* ut, Real* Aj, Real* Bj, Real* Cj) {
  const Real dt = P.dt;
  if (P.dynamics == RDA_DYN_OMNI) {
    Real phi = ut[1], v = ut[0];
    Real c = cos(phi), s = sin(phi);
    Aj[0] = 0; Aj[1] = 0;
    Bj[0] = c * dt; Bj[1] = -v * s * dt; Bj[2] = s * dt; Bj[3] = v * c * dt; Bj[4] = 0; Bj[5] = 0;
    Cj[0] = phi * v * s * dt; Cj[1] = -phi * v * c * dt; Cj[2] = 0;
    return;
  }
  Real phi = st[2], v = ut[0];
  Real c = cos(phi), s = sin(phi);
  Aj[0] = -v * dt * s; Aj[1] = v * dt * c;
  Bj[0] = c * dt; Bj[1] = 0; Bj[2] = s * dt; Bj[3] = 0;
  Cj[0] = phi * v * s * dt; Cj[1] = -phi * v * c * dt;
  if (P.dynamics == RDA_DYN_ACKER) {
    Real psi = ut[1];
    Real cp = cos(psi);
    Real k = v * dt / ((Real)P.L * cp * cp);
    Bj[4] = tan(psi) * dt / (Real)P.L; Bj[5] = k;
    Cj[2] = -psi * k;
  } else {
    Bj[4] = 0; Bj[5] = dt;
    Cj[2] = 0;
  }
}

// One inequality row of stage t.  c in 0..9: (u0 hi, u0 lo, u1 hi, u1 lo, d hi, d lo,
// rate0 hi, rate0 lo, rate1 hi, rate1 lo).  value g >= 0, gradient = sgn on component `comp`
// (3: u0, 4: u1, 5: d) and -sgn on the previous control for rate rows.
template <typename Real>
struct Row { Real g; Real sgn; int comp; bool rate; bool live; };

template <typename Real>
RDA_HD Row<Real> su_row(const SuParams& P, const SuWork<Real>& W, int t, int c) {
  Row<Real> r;
  r.rate = c >= 6;
  r.live = true;
  const bool hi = (c & 1) == 0;
  r.sgn = hi ? (Real)-1 : (Real)1;
  if (c < 4) {
    int k = c >> 1;
    r.comp = 3 + k;
    Real uv = W.u[2 * t + k], m = P.umax[k];
    r.g = hi ? m - uv : uv + m;
  } else if (c < 6) {
    r.comp = 5;
    Real dv = W.d[t];
    Real lo = P.dmin > 0 ? P.dmin : 0;
    r.g = hi ? (Real)P.dmax - dv : dv - lo;
    r.live = P.N > 0;
  } else {
    int k = (c - 6) >> 1;
    r.comp = 3 + k;
    r.live = t >= 1;
    Real du = r.live ? W.u[2 * t + k] - W.u[2 * (t - 1) + k] : (Real)0;
    r.g = hi ? (Real)P.ab[k] - du : (Real)P.ab[k] + du;
  }
  return r;
}

// directional derivative of row (t, c) along the Newton step held in (dz, dv)
template <typename Real>
RDA_HD Real su_row_dir(const Row<Real>& r, const Real* dz_t, const Real* dv_t) {
  Real x = dv_t[r.comp - 3];
  if (r.rate) x -= dz_t[3 + (r.comp - 3)];
  return r.sgn * x;
}

template <typename Real, typename Ctx>
RDA_HD void su_riccati(const SuParams& P, SuWork<Real>& W, Ctx& ctx, bool factor, Real* dz, Real* dv) {
  // Riccati recursion of the Newton step (banded KKT system).  Stage t: state z = (s_t, u_{t-1}),
  // control v = (u_t, d_t); the stage cost is quadratic in q = (s_{t+1}, u_t, d_t) = J [z; v] with
  // J = [[A 0 B 0], [0 0 I 0], [0 0 0 1]] plus the rate-limit coupling between u_t and u_{t-1}.
  // Factor pass: the matrix work of a stage is spread over the lanes entry by entry (Q = P+ + stage
  // terms, T = Q J, H = J'T, K = -Hvv^-1 Hvz, P = Hzz + Hvz'K) with the blocks staged in shared
  // memory (W.rw); the small vector recursion (cost-to-go gradient) is carried redundantly in
  // registers by every lane.
  const int T = P.T;
  const int lane = ctx.lane(), nl = ctx.nlanes();
  const Real reg = (Real)1e-9;
  const Real tw = 2 * (Real)P.ws, tw3 = (P.dynamics == RDA_DYN_OMNI ? (Real)0 : tw);
  Real* Qs = W.rw;          // 36  Q[a*6+b]
  Real* Ts = W.rw + 36;     // 18  t2[r], t5[r], t6[r]  (r = 0..5)
  Real* Hz = W.rw + 54;     // 6   Hzz dense part: (0,0) (0,1) (1,1) (0,2) (1,2) (2,2)
  Real* Hv = W.rw + 60;     // 15  Hvz[k*5+b]
  Real* Hh = W.rw + 75;     // 6   Hvv: h00 h10 h11 h20 h21 h22
  Real* Pm = W.rw + 81;     // 25  cost-to-go Hessian P[a*5+b]
  Real pv[5] = {0, 0, 0, 0, 0};
  const bool writer = lane == 0;
  if (factor) {
    for (int e = lane; e < 25; e += nl) Pm[e] = 0;
    ctx.sync();
  }
  for (int t = T - 1; t >= 0; --t) {
    const Real a02 = W.Aj[2 * t], a12 = W.Aj[2 * t + 1];
    const Real* Bt = W.Bj + 6 * t;
    const Real b00 = Bt[0], b01 = Bt[1], b10 = Bt[2], b11 = Bt[3], b20 = Bt[4], b21 = Bt[5];
    const Real* wb = W.wb + 5 * t;
    const Real wr0 = wb[3], wr1 = wb[4];
    // gradient in q-space (+ cost-to-go), pulled back through J
    const Real* gw = W.gw + 8 * t;
    const Real q0 = gw[0] + pv[0], q1 = gw[1] + pv[1], q2 = gw[2] + pv[2], q3 = gw[3] + pv[3],
               q4 = gw[4] + pv[4], q5 = gw[5];
    const Real gz0 = q0, gz1 = q1, gz2 = a02 * q0 + a12 * q1 + q2, gz3 = gw[6], gz4 = gw[7];
    const Real gv0 = b00 * q0 + b10 * q1 + b20 * q2 + q3;
    const Real gv1 = b01 * q0 + b11 * q1 + b21 * q2 + q4;
    const Real gv2 = q5;
    Real i00, L10, i11, L20, L21, i22;   // Cholesky of Hvv, reciprocal diagonal
    Real* Kt = W.K + 15 * t;             // K[k*5+b]
    if (factor) {
      const Real* M = W.Wm + 6 * t;
      // S1: Q = P+ (5x5, zero-padded) + stage terms
      for (int e = lane; e < 36; e += nl) {
        const int a = e / 6, b = e - 6 * a;
        Real v = (a < 5 && b < 5) ? Pm[a * 5 + b] : (Real)0;
        if (a == b) {
          if (a == 0) v += tw + M[0];
          else if (a == 1) v += tw + M[3];
          else if (a == 2) v += tw3 + (Real)P.ro2 * W.Skk[t];
          else if (a == 3) v += 2 * (Real)P.wu + reg + wb[0] + wr0;
          else if (a == 4) v += reg + wb[1] + wr1;
          else v = (P.N > 0 ? reg + wb[2] + M[5] : (Real)1);
        } else {
          const int lo = a < b ? a : b, hi = a < b ? b : a;
          if (lo == 0 && hi == 1) v += M[1];
          else if (lo == 0 && hi == 5) v = M[2];
          else if (lo == 1 && hi == 5) v = M[4];
        }
        Qs[e] = v;
      }
      ctx.sync();
      // S2: T = Q J for the three columns of J that are not unit vectors
      for (int e = lane; e < 18; e += nl) {
        const int c = e / 6, r = e - 6 * c;
        const Real* Qr = Qs + 6 * r;
        Real v;
        if (c == 0) v = a02 * Qr[0] + a12 * Qr[1] + Qr[2];
        else if (c == 1) v = b00 * Qr[0] + b10 * Qr[1] + b20 * Qr[2] + Qr[3];
        else v = b01 * Qr[0] + b11 * Qr[1] + b21 * Qr[2] + Qr[4];
        Ts[e] = v;
      }
      ctx.sync();
      // S3: H = J'T (27 entries: 6 of Hzz, 15 of Hvz, 6 of Hvv)
      for (int e = lane; e < 27; e += nl) {
        const Real* t2 = Ts; const Real* t5 = Ts + 6; const Real* t6 = Ts + 12;
#define RDA_J2(x) (a02 * (x)[0] + a12 * (x)[1] + (x)[2])
#define RDA_J5(x) (b00 * (x)[0] + b10 * (x)[1] + b20 * (x)[2] + (x)[3])
#define RDA_J6(x) (b01 * (x)[0] + b11 * (x)[1] + b21 * (x)[2] + (x)[4])
        Real v;
        switch (e) {
          case 0: v = Qs[0]; break;               // Hzz(0,0)
          case 1: v = Qs[1]; break;               // Hzz(0,1)
          case 2: v = Qs[7]; break;               // Hzz(1,1)
          case 3: v = t2[0]; break;               // Hzz(0,2)
          case 4: v = t2[1]; break;               // Hzz(1,2)
          case 5: v = RDA_J2(t2); break;          // Hzz(2,2)
          case 6: v = t5[0]; break;               // Hvz(0,0..4)
          case 7: v = t5[1]; break;
          case 8: v = RDA_J2(t5); break;
          case 9: v = -wr0; break;
          case 10: v = 0; break;
          case 11: v = t6[0]; break;              // Hvz(1,0..4)
          case 12: v = t6[1]; break;
          case 13: v = RDA_J2(t6); break;
          case 14: v = 0; break;
          case 15: v = -wr1; break;
          case 16: v = Qs[30]; break;             // Hvz(2,0..4): row 5 of Q
          case 17: v = Qs[31]; break;
          case 18: v = RDA_J2(Qs + 30); break;
          case 19: v = 0; break;
          case 20: v = 0; break;
          case 21: v = RDA_J5(t5); break;         // h00
          case 22: v = RDA_J6(t5); break;         // h10
          case 23: v = RDA_J6(t6); break;         // h11
          case 24: v = t5[5]; break;              // h20
          case 25: v = t6[5]; break;              // h21
          default: v = Qs[35]; break;             // h22
        }
#undef RDA_J2
#undef RDA_J5
#undef RDA_J6
        if (e < 6) Hz[e] = v;
        else if (e < 21) Hv[e - 6] = v;
        else Hh[e - 21] = v;
      }
      ctx.sync();
      // S4: Cholesky of Hvv (every lane, registers)
      const Real L00 = sqrt_(Hh[0]);
      i00 = (Real)1 / L00;
      L10 = Hh[1] * i00;
      const Real L11 = sqrt_(Hh[2] - L10 * L10);
      i11 = (Real)1 / L11;
      L20 = Hh[3] * i00;
      L21 = (Hh[4] - L20 * L10) * i11;
      const Real L22 = sqrt_(Hh[5] - L20 * L20 - L21 * L21);
      i22 = (Real)1 / L22;
      // S5: K = -Hvv^-1 Hvz, one column per lane
      for (int b = lane; b < 5; b += nl) {
        Real y0 = -Hv[b] * i00;
        Real y1 = (-Hv[5 + b] - L10 * y0) * i11;
        Real y2 = (-Hv[10 + b] - L20 * y0 - L21 * y1) * i22;
        Real x2 = y2 * i22;
        Real x1 = (y1 - L21 * x2) * i11;
        Real x0 = (y0 - L10 * x1 - L20 * x2) * i00;
        Kt[b] = x0; Kt[5 + b] = x1; Kt[10 + b] = x2;
      }
      if (writer) {
        Real* Ls = W.Lc + 6 * t;
        Ls[0] = i00; Ls[1] = L10; Ls[2] = i11; Ls[3] = L20; Ls[4] = L21; Ls[5] = i22;
      }
      ctx.sync();
      // S6: P = Hzz + Hvz'K (25 entries)
      for (int e = lane; e < 25; e += nl) {
        const int a = e / 5, b = e - 5 * a;
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        Real hzz = 0;
        if (hi < 3) hzz = Hz[lo == 0 ? (hi == 0 ? 0 : hi == 1 ? 1 : 3) : lo == 1 ? (hi == 1 ? 2 : 4) : 5];
        else if (lo == hi) hzz = (hi == 3 ? wr0 : wr1);
        // symmetric evaluation: use (lo, hi) ordering so both triangles get identical values
        Pm[e] = hzz + Hv[lo] * Kt[hi] + Hv[5 + lo] * Kt[5 + hi] + Hv[10 + lo] * Kt[10 + hi];
      }
      ctx.sync();
    } else {
      const Real* Ls = W.Lc + 6 * t;
      i00 = Ls[0]; L10 = Ls[1]; i11 = Ls[2]; L20 = Ls[3]; L21 = Ls[4]; i22 = Ls[5];
    }
    {
      Real y0 = -gv0 * i00;
      Real y1 = (-gv1 - L10 * y0) * i11;
      Real y2 = (-gv2 - L20 * y0 - L21 * y1) * i22;
      Real x2 = y2 * i22;
      Real x1 = (y1 - L21 * x2) * i11;
      Real x0 = (y0 - L10 * x1 - L20 * x2) * i00;
      if (writer) { W.kf[3 * t] = x0; W.kf[3 * t + 1] = x1; W.kf[3 * t + 2] = x2; }
      pv[0] = gz0 + Kt[0] * gv0 + Kt[5] * gv1 + Kt[10] * gv2;
      pv[1] = gz1 + Kt[1] * gv0 + Kt[6] * gv1 + Kt[11] * gv2;
      pv[2] = gz2 + Kt[2] * gv0 + Kt[7] * gv1 + Kt[12] * gv2;
      pv[3] = gz3 + Kt[3] * gv0 + Kt[8] * gv1 + Kt[13] * gv2;
      pv[4] = gz4 + Kt[4] * gv0 + Kt[9] * gv1 + Kt[14] * gv2;
    }
  }
  ctx.sync();
  // forward sweep
  Real z[5] = {0, 0, 0, 0, 0};
  for (int t = 0; t < T; ++t) {
    Real v[3];
    for (int k = 0; k < 3; ++k) {
      Real sacc = W.kf[3 * t + k];
      for (int b = 0; b < 5; ++b) sacc += W.K[15 * t + 5 * k + b] * z[b];
      v[k] = sacc;
    }
    if (writer) {
      for (int a = 0; a < 5; ++a) dz[5 * t + a] = z[a];
      for (int k = 0; k < 3; ++k) dv[3 * t + k] = v[k];
    }
    Real n0 = z[0] + W.Aj[2 * t] * z[2] + W.Bj[6 * t] * v[0] + W.Bj[6 * t + 1] * v[1];
    Real n1 = z[1] + W.Aj[2 * t + 1] * z[2] + W.Bj[6 * t + 2] * v[0] + W.Bj[6 * t + 3] * v[1];
    Real n2 = z[2] + W.Bj[6 * t + 4] * v[0] + W.Bj[6 * t + 5] * v[1];
    z[0] = n0; z[1] = n1; z[2] = n2; z[3] = v[0]; z[4] = v[1];
  }
  if (writer) for (int a = 0; a < 5; ++a) dz[5 * T + a] = z[a];
  ctx.sync();
}

// Solve the su-QP.  Inputs already staged in W: lins, linu, ref, vref, hx/hy/hc, pref, Skk/Sgk
// are computed here from (gx, gy) planes passed as pointers (global or shared memory, [o*T+t]).
// On entry W.d holds para_dis (initial guess of d).  Returns 0 (converged), 1 (iteration cap),
// 2 (non-finite).  On return W.s, W.u, W.d hold the solution.
template <typename Real, typename Ctx>
RDA_HD int su_solve(const SuParams& P, SuWork<Real>& W, Ctx& ctx, const float* gx, const float* gy,
                    int* iters_out) {
  const int T = P.T, N = P.N;
  const int lane = ctx.lane(), nl = ctx.nlanes();
  const Real ro1 = P.ro1, ro2 = P.ro2;
  const Real iro1 = (Real)1 / ro1;
  const bool acc = P.accelerated != 0;
  // ---- linearisation and aggregated rotation terms (lane-parallel over stages) ----
  for (int t = lane; t < T; t += nl) {
    su_linearise<Real>(P, W.lins + 3 * t, W.linu + 2 * t, W.Aj + 2 * t, W.Bj + 6 * t, W.Cj + 3 * t);
    Real phib = W.lins[3 * t + 2];
    Real c = cos(phib), s = sin(phib);
    W.cph[t] = c; W.sph[t] = s;
    Real skk = 0, sgk = 0;
    for (int o = 0; o < N; ++o) {
      Real ax = W.hx[o * T + t], ay = W.hy[o * T + t];
      Real k0 = -ax * s + ay * c, k1 = -ax * c - ay * s;          // a R'
      Real g0 = (Real)gx[o * T + t] + ax * c + ay * s;            // mu'G + xi + a R
      Real g1 = (Real)gy[o * T + t] - ax * s + ay * c;
      skk += k0 * k0 + k1 * k1;
      sgk += g0 * k0 + g1 * k1;
    }
    W.Skk[t] = skk; W.Sgk[t] = sgk;
    W.u[2 * t] = W.linu[2 * t]; W.u[2 * t + 1] = W.linu[2 * t + 1];
  }
  ctx.sync();
  // ---- initial iterate: roll the linearised model out from s_0 ----
  if (lane == 0) {
    W.s[0] = W.lins[0]; W.s[1] = W.lins[1]; W.s[2] = W.lins[2];
    for (int t = 0; t < T; ++t) {
      const Real* s0 = W.s + 3 * t;
      Real u0 = W.u[2 * t], u1 = W.u[2 * t + 1];
      W.s[3 * t + 3] = s0[0] + W.Aj[2 * t] * s0[2] + W.Bj[6 * t] * u0 + W.Bj[6 * t + 1] * u1 + W.Cj[3 * t];
      W.s[3 * t + 4] = s0[1] + W.Aj[2 * t + 1] * s0[2] + W.Bj[6 * t + 2] * u0 + W.Bj[6 * t + 3] * u1 + W.Cj[3 * t + 1];
      W.s[3 * t + 5] = s0[2] + W.Bj[6 * t + 4] * u0 + W.Bj[6 * t + 5] * u1 + W.Cj[3 * t + 2];
    }
  }
  ctx.sync();
  const Real mu0 = P.mu0 > 0 ? (Real)P.mu0 : (Real)1;
  int nrows = 0;
  for (int t = lane; t < T; t += nl) {
    for (int c = 0; c < 10; ++c) {
      Row<Real> r = su_row<Real>(P, W, t, c);
      Real sv = r.live ? rmax(r.g, (Real)1e-2) : (Real)1;
      W.bs[10 * t + c] = sv;
      W.bnu[10 * t + c] = r.live ? mu0 / sv : (Real)0;
      if (r.live) ++nrows;
    }
    if (acc) {
      Real dx = W.s[3 * t + 3] - W.pref[2 * t], dy = W.s[3 * t + 4] - W.pref[2 * t + 1];
      for (int o = 0; o < N; ++o) {
        Real l = (Real)W.hx[o * T + t] * dx + (Real)W.hy[o * T + t] * dy + (Real)W.hc[o * T + t] - W.d[t];
        Real sv = (l + sqrt_(l * l + 4 * mu0 / ro1)) / 2;
        W.hs[o * T + t] = sv;
        W.hnu[o * T + t] = mu0 / sv;
        ++nrows;
      }
    }
  }
  const Real Mrows = ctx.sum((Real)nrows);
  ctx.sync();
  const Real tol_mu = sizeof(Real) == 4 ? (Real)1e-6 : (Real)1e-10;
  const Real tol_r = sizeof(Real) == 4 ? (Real)1e-5 : (Real)1e-9;
  const Real reg = (Real)1e-9;
  const Real tol_step = sizeof(Real) == 4 ? (Real)2e-4 : (Real)1e-7;
  const Real tol_floor = sizeof(Real) == 4 ? (Real)1e-7 : (Real)1e-13;
  Real last_step = 1e30f;       // size of the previous Newton update (stationarity proxy)
  int status = 1, it = 0;
  for (it = 0; it < P.max_iter; ++it) {
    Real sigma_mu = 0;
    Real mu = 0;
    for (int phase = 0; phase < 2; ++phase) {
      // ---- assemble stage gradients (and, in phase 0, Hessian weights) ----
      Real acc_mu = 0, acc_r = 0;
      for (int t = lane; t < T; t += nl) {
        Real* gw = W.gw + 8 * t;
        const Real* sn = W.s + 3 * t + 3;
        const Real tw = 2 * (Real)P.ws;
        gw[0] = tw * (sn[0] - W.ref[3 * t + 3]);
        gw[1] = tw * (sn[1] - W.ref[3 * t + 4]);
        gw[2] = (P.dynamics == RDA_DYN_OMNI ? (Real)0 : tw * (sn[2] - W.ref[3 * t + 5]))
                + ro2 * (W.Skk[t] * (sn[2] - W.lins[3 * t + 2]) + W.Sgk[t]);
        gw[3] = 2 * (Real)P.wu * (W.u[2 * t] - W.vref) + reg * W.u[2 * t];
        gw[4] = reg * W.u[2 * t + 1];
        gw[5] = N > 0 ? -(Real)P.slack_gain + reg * W.d[t] : (Real)0;
        gw[6] = 0; gw[7] = 0;
        Real wb[5] = {0, 0, 0, 0, 0};
        for (int c = 0; c < 10; ++c) {
          Row<Real> r = su_row<Real>(P, W, t, c);
          if (!r.live) continue;
          Real sv = W.bs[10 * t + c], nu = W.bnu[10 * t + c];
          Real res = r.g - sv;
          const Real isv = (Real)1 / sv;
          Real om = nu * isv;
          Real term;
          if (phase == 0) {
            term = -om * res;
            acc_mu += sv * nu;
            acc_r = rmax(acc_r, abs_(res));
          } else {
            Real dir = su_row_dir<Real>(r, W.dza + 5 * t, W.dva + 3 * t);
            Real dsa = dir + res;
            Real dna = -nu - om * dsa;
            term = (sigma_mu - dsa * dna) * isv - om * res;
          }
          // g -= grad * term
          gw[r.comp] -= r.sgn * term;
          if (r.rate) gw[r.comp + 3] += r.sgn * term;
          if (phase == 0) wb[(r.rate ? 3 : 0) + (r.comp - 3)] += om;
        }
        if (phase == 0) for (int k = 0; k < 5; ++k) W.wb[5 * t + k] = wb[k];
        Real m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0;
        Real dx = sn[0] - W.pref[2 * t], dy = sn[1] - W.pref[2 * t + 1];
        Real dd = W.d[t];
        for (int o = 0; o < N; ++o) {
          Real ax = W.hx[o * T + t], ay = W.hy[o * T + t];
          Real l = ax * dx + ay * dy + (Real)W.hc[o * T + t] - dd;
          Real tk, om;
          if (acc) {
            Real sv = W.hs[o * T + t], nu = W.hnu[o * T + t];
            const Real nr = nu * iro1;
            Real res = l + nr - sv;
            const Real iden = (Real)1 / (sv + nr);
            om = nu * iden;
            if (phase == 0) {
              tk = om * (nr - res);
              acc_mu += sv * nu;
              acc_r = rmax(acc_r, abs_(res));
            } else {
              Real dir = ax * W.dza[5 * t + 5] + ay * W.dza[5 * t + 6] - W.dva[3 * t + 2];
              Real dna = -om * (sv + res + dir);
              Real dsa = dir + dna * iro1 + res;
              Real cc = sv * nu - sigma_mu + dsa * dna;
              tk = nu - (cc + nu * res) * iden;
            }
          } else {
            om = ro1;               // plain quadratic 1/2 ro1 Im^2  (rda_solver.py:378-379)
            tk = -ro1 * l;
          }
          gw[0] -= ax * tk; gw[1] -= ay * tk; gw[5] += tk;
          if (phase == 0) {
            m0 += om * ax * ax; m1 += om * ax * ay; m2 -= om * ax;
            m3 += om * ay * ay; m4 -= om * ay; m5 += om;
          }
        }
        if (phase == 0) {
          Real* M = W.Wm + 6 * t;
          M[0] = m0; M[1] = m1; M[2] = m2; M[3] = m3; M[4] = m4; M[5] = m5;
        }
      }
      if (phase == 0) {
        mu = ctx.sum(acc_mu) / Mrows;
        Real rmx = ctx.max(acc_r);
#ifdef RDA_SU_DEBUG
        printf("it %d mu %.3e rmx %.3e last_step %.3e\n", it, (double)mu, (double)rmx, (double)last_step);
#endif
        if (!finite_(mu)) { status = 2; break; }
        // converged: complementarity and residuals small, and the last Newton update small
        // (degenerate problems approach the solution like sqrt(mu): stop at the rounding floor)
        if (mu < tol_mu && rmx < tol_r && (last_step < tol_step || mu < tol_floor)) { status = 0; break; }
      }
      ctx.sync();
      su_riccati<Real, Ctx>(P, W, ctx, phase == 0, phase == 0 ? W.dza : W.dz, phase == 0 ? W.dva : W.dv);
      // ---- step lengths ----
      const Real* dz = phase == 0 ? W.dza : W.dz;
      const Real* dv = phase == 0 ? W.dva : W.dv;
      Real rmaxr = 0, s0 = 0, s1 = 0, s2 = 0;   // rmaxr = max over rows of (-delta / value): 1 / max step
      for (int t = lane; t < T; t += nl) {
        for (int c = 0; c < 10; ++c) {
          Row<Real> r = su_row<Real>(P, W, t, c);
          if (!r.live) continue;
          Real sv = W.bs[10 * t + c], nu = W.bnu[10 * t + c];
          Real res = r.g - sv;
          Real dir = su_row_dir<Real>(r, dz + 5 * t, dv + 3 * t);
          Real ds = dir + res, dn;
          const Real ip = (Real)1 / (sv * nu);
          const Real isv = nu * ip, inu = sv * ip, om = nu * isv;
          if (phase == 0) dn = -nu - om * ds;
          else {
            Real dira = su_row_dir<Real>(r, W.dza + 5 * t, W.dva + 3 * t);
            Real dsa = dira + res;
            Real dna = -nu - om * dsa;
            dn = (sigma_mu - dsa * dna) * isv - nu - om * ds;
          }
          rmaxr = rmax(rmaxr, rmax(-ds * isv, -dn * inu));
          s0 += sv * nu; s1 += sv * dn + nu * ds; s2 += ds * dn;
        }
        if (acc) {
          Real dx = W.s[3 * t + 3] - W.pref[2 * t], dy = W.s[3 * t + 4] - W.pref[2 * t + 1];
          for (int o = 0; o < N; ++o) {
            Real ax = W.hx[o * T + t], ay = W.hy[o * T + t];
            Real l = ax * dx + ay * dy + (Real)W.hc[o * T + t] - W.d[t];
            Real sv = W.hs[o * T + t], nu = W.hnu[o * T + t];
            const Real nr = nu * iro1;
            Real res = l + nr - sv;
            const Real iden = (Real)1 / (sv + nr);
            const Real om = nu * iden;
            Real dir = ax * dz[5 * t + 5] + ay * dz[5 * t + 6] - dv[3 * t + 2];
            Real cc = sv * nu;
            if (phase == 1) {
              Real dira = ax * W.dza[5 * t + 5] + ay * W.dza[5 * t + 6] - W.dva[3 * t + 2];
              Real dna = -om * (sv + res + dira);
              Real dsa = dira + dna * iro1 + res;
              cc = sv * nu - sigma_mu + dsa * dna;
            }
            Real dn = -(cc + nu * res + nu * dir) * iden;
            Real ds = dir + dn * iro1 + res;
            const Real ip = (Real)1 / (sv * nu);
            rmaxr = rmax(rmaxr, rmax(-ds * nu * ip, -dn * sv * ip));
            s0 += sv * nu; s1 += sv * dn + nu * ds; s2 += ds * dn;
          }
        }
      }
      rmaxr = ctx.max(rmaxr);
      const Real amax = rmaxr > (Real)1e-30 ? (Real)1 / rmaxr : (Real)1e30;
      if (phase == 0) {
        Real a = rmin((Real)1, amax);
        Real mua = (ctx.sum(s0) + a * ctx.sum(s1) + a * a * ctx.sum(s2)) / Mrows;
        Real sg = mua / mu;
        sg = sg * sg * sg;
        sigma_mu = rmin(sg, (Real)1) * mu;
      } else {
        Real a = rmin((Real)1, (Real)0.995 * amax);
#ifdef RDA_SU_DEBUG
        printf("   alpha %.3e sigma_mu %.3e\n", (double)a, (double)sigma_mu);
#endif
        // ---- update (needs the corrector quantities once more) ----
        for (int t = lane; t < T; t += nl) {
          // rows first: they read the OLD iterate through su_row
          Real gsave[10];
          for (int c = 0; c < 10; ++c) { Row<Real> r = su_row<Real>(P, W, t, c); gsave[c] = r.g; }
          for (int c = 0; c < 10; ++c) {
            Row<Real> r = su_row<Real>(P, W, t, c);
            if (!r.live) continue;
            Real sv = W.bs[10 * t + c], nu = W.bnu[10 * t + c];
            Real res = gsave[c] - sv;
            Real dir = su_row_dir<Real>(r, W.dz + 5 * t, W.dv + 3 * t);
            Real ds = dir + res;
            Real dira = su_row_dir<Real>(r, W.dza + 5 * t, W.dva + 3 * t);
            Real dsa = dira + res;
            const Real isv = (Real)1 / sv, om = nu * isv;
            Real dna = -nu - om * dsa;
            Real dn = (sigma_mu - dsa * dna) * isv - nu - om * ds;
            W.bs[10 * t + c] = sv + a * ds;
            W.bnu[10 * t + c] = nu + a * dn;
          }
          if (acc) {
            Real dx = W.s[3 * t + 3] - W.pref[2 * t], dy = W.s[3 * t + 4] - W.pref[2 * t + 1];
            for (int o = 0; o < N; ++o) {
              Real ax = W.hx[o * T + t], ay = W.hy[o * T + t];
              Real l = ax * dx + ay * dy + (Real)W.hc[o * T + t] - W.d[t];
              Real sv = W.hs[o * T + t], nu = W.hnu[o * T + t];
              const Real nr = nu * iro1;
              Real res = l + nr - sv;
              const Real iden = (Real)1 / (sv + nr);
              const Real om = nu * iden;
              Real dir = ax * W.dz[5 * t + 5] + ay * W.dz[5 * t + 6] - W.dv[3 * t + 2];
              Real dira = ax * W.dza[5 * t + 5] + ay * W.dza[5 * t + 6] - W.dva[3 * t + 2];
              Real dna = -om * (sv + res + dira);
              Real dsa = dira + dna * iro1 + res;
              Real cc = sv * nu - sigma_mu + dsa * dna;
              Real dn = -(cc + nu * res + nu * dir) * iden;
              Real ds = dir + dn * iro1 + res;
              W.hs[o * T + t] = sv + a * ds;
              W.hnu[o * T + t] = nu + a * dn;
            }
          }
        }
        ctx.sync();   // every lane has finished reading the old (s, u, d) of its neighbours
        Real stepmax = 0;
        for (int t = lane; t < T; t += nl) {
          stepmax = rmax(stepmax, rmax(abs_(W.dv[3 * t]), rmax(abs_(W.dv[3 * t + 1]), abs_(W.dv[3 * t + 2]))));
          W.u[2 * t] += a * W.dv[3 * t]; W.u[2 * t + 1] += a * W.dv[3 * t + 1];
          if (N > 0) W.d[t] += a * W.dv[3 * t + 2];
          W.s[3 * t + 3] += a * W.dz[5 * t + 5];
          W.s[3 * t + 4] += a * W.dz[5 * t + 6];
          W.s[3 * t + 5] += a * W.dz[5 * t + 7];
        }
        last_step = a * ctx.max(stepmax);
        ctx.sync();
      }
    }
    if (status != 1) break;
  }
  if (iters_out) *iters_out = it;
  return status;
}

}  // namespace rda
