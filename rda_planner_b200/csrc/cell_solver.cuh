// cell_solver.cuh — one (obstacle, stage) cell of the LamMuZ problem and its multiplier update.
//
// Replaces, per cell, the reference's per-obstacle cvxpy/ECOS solve and the numpy update loops:
//   problem      rda_solver.py:389-421 (LamMuZ_cost_cons), :874-909 (Hm_LamMu, Im_LamMu),
//                cones :1034-1050; separable over the horizon (max of norms <= 1, :408-416)
//   updates      update_xi :668-690, update_zeta :639-666, lam'A / lam'b :529-542
//
// Method.  With v = A'lam, g = G'mu and support functions sigma_O, sigma_Rob the
// cell is the 4-variable convex problem
//     min_{|v|<=1, g}  1/2 neg(stuff)^2 + ro2/2 |g + R'v + xi|^2,
//     stuff = v.p - sigma_O(v) - sigma_Rob(g) - d + zeta,
// solved in coordinates relative to the robot reference point p:
//   fast path : xi == 0 and obstacle/robot disjoint  -> closest pair of points (exact);
//               robot-vertex contact candidates accepted by their KKT conditions (exact);
//   slow path : small dense primal-dual interior point method in float64.
// Tie-break where the reference argmin is not unique: max margin, LP-vertex multipliers,
// z = theta * max(stuff, 0).
#pragma once
#include "rda_hd.h"

namespace rda {

enum { CELL_FAST_INACTIVE = 0, CELL_FAST_VERTEX = 1, CELL_SLOW_A = 2, CELL_SLOW_B = 3,
       CELL_OVERLAP_FREE = 4, CELL_FAILED = 5, CELL_NEEDS_SLOW = 6 };

template <typename Real>
struct CellOut {
  Real lam[RDA_MAX_EDGE];
  Real mu[RDA_MAX_ROBOT_EDGE];
  Real z, zeta_new, xi0_new, xi1_new;   // multiplier state after the update
  Real ax, ay, c0, gx, gy;              // su-QP hinge inputs (lam'A, offset, mu'G + xi)
  Real hm0, hm1;                        // Hm of this cell (primal residual, :682)
  int path;
};

// ---------------------------------------------------------------------------------------------
// Small dense primal-dual interior point method:
//    min 1/2 x'Qx + c'x   s.t.  a_i'x <= b_i (i < m),  x0^2 + x1^2 <= 1
// NV <= 7 variables, m <= 2*8+1 rows.  Mehrotra predictor-corrector, float64.
// ---------------------------------------------------------------------------------------------
template <int NV, int MC>
struct TinyQP {
  double Q[NV][NV];
  double c[NV];
  // sparse rows a_i'x <= b_i: at most 4 non-zeros each
  double av[MC][4];
  int ai[MC][4];
  int an[MC];
  double b[MC];
  int m;
  RDA_HD void row(int n, int i0, double v0, int i1, double v1, int i2, double v2, int i3, double v3, double rhs) {
    av[m][0] = v0; av[m][1] = v1; av[m][2] = v2; av[m][3] = v3;
    ai[m][0] = i0; ai[m][1] = i1; ai[m][2] = i2; ai[m][3] = i3;
    an[m] = n; b[m] = rhs; ++m;
  }
  RDA_HD double dot(int i, const double* x) const {
    double sacc = 0;
    for (int e = 0; e < an[i]; ++e) sacc += av[i][e] * x[ai[i][e]];
    return sacc;
  }
  RDA_HD void axpy(int i, double w, double* y) const {      // y += w * a_i
    for (int e = 0; e < an[i]; ++e) y[ai[i][e]] += w * av[i][e];
  }
  RDA_HD void rank1(int i, double w, double H[NV][NV]) const {   // lower triangle of H += w a_i a_i'
    for (int e = 0; e < an[i]; ++e)
      for (int f = 0; f < an[i]; ++f) {
        int r = ai[i][e], cidx = ai[i][f];
        if (r >= cidx) H[r][cidx] += w * av[i][e] * av[i][f];
      }
  }
  int tv;   // -1: unit-ball constraint x0^2 + x1^2 <= 1;  k >= 0: cone constraint
            // (x0^2 + x1^2)/x_k - x_k <= 0 (i.e. |x01| <= x_k, rows keep 0 <= x_k <= 1)
};

template <int NV>
RDA_HD bool chol_solve(double H[NV][NV], double* r1, double* r2) {
  // in-place Cholesky H = LL', then solve for two right-hand sides
  for (int j = 0; j < NV; ++j) {
    double d = H[j][j];
    for (int k = 0; k < j; ++k) d -= H[j][k] * H[j][k];
    if (!(d > 0)) return false;
    d = sqrt(d);
    H[j][j] = d;
    for (int i = j + 1; i < NV; ++i) {
      double s = H[i][j];
      for (int k = 0; k < j; ++k) s -= H[i][k] * H[j][k];
      H[i][j] = s / d;
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    double* r = pass ? r2 : r1;
    if (!r) continue;
    for (int i = 0; i < NV; ++i) {
      double s = r[i];
      for (int k = 0; k < i; ++k) s -= H[i][k] * r[k];
      r[i] = s / H[i][i];
    }
    for (int i = NV - 1; i >= 0; --i) {
      double s = r[i];
      for (int k = i + 1; k < NV; ++k) s -= H[k][i] * r[k];
      r[i] = s / H[i][i];
    }
  }
  return true;
}

// value, gradient (on x0, x1 and x_tv) and Hessian of the one nonlinear constraint
struct ConeEval { double f, g0, g1, gt, h00, h11, h0t, h1t, htt; };
RDA_HD ConeEval cone_eval(const double* x, int tv) {
  ConeEval c;
  if (tv < 0) {
    c.f = x[0] * x[0] + x[1] * x[1] - 1.0;
    c.g0 = 2 * x[0]; c.g1 = 2 * x[1]; c.gt = 0;
    c.h00 = 2; c.h11 = 2; c.h0t = 0; c.h1t = 0; c.htt = 0;
  } else {
    double t = x[tv], n2 = x[0] * x[0] + x[1] * x[1];
    c.f = n2 / t - t;
    c.g0 = 2 * x[0] / t; c.g1 = 2 * x[1] / t; c.gt = -n2 / (t * t) - 1.0;
    c.h00 = 2 / t; c.h11 = 2 / t; c.h0t = -2 * x[0] / (t * t); c.h1t = -2 * x[1] / (t * t);
    c.htt = 2 * n2 / (t * t * t);
  }
  return c;
}

template <int NV, int MC>
RDA_HD_NOINLINE bool tiny_ipm(const TinyQP<NV, MC>& P, double* x /* in: strictly feasible start */) {
  const int m = P.m;
  double s[MC + 1], l[MC + 1];
  for (int i = 0; i < m; ++i) {
    s[i] = rmax(P.b[i] - P.dot(i, x), 1e-3);
    l[i] = 1.0 / s[i];
  }
  const int tv = P.tv;
  s[m] = rmax(-cone_eval(x, tv).f, 1e-3);
  l[m] = 1.0 / s[m];
  const int M = m + 1;
  double scale = 1.0;
  for (int i = 0; i < m; ++i) {
    scale = rmax(scale, fabs(P.b[i]));
    for (int e = 0; e < P.an[i]; ++e) scale = rmax(scale, fabs(P.av[i][e]));
  }
  for (int k = 0; k < NV; ++k) scale = rmax(scale, fabs(P.c[k]));
  bool acceptable = false;
  for (int it = 0; it < 40; ++it) {
    // residuals
    double rd[NV], rp[MC + 1];
    for (int k = 0; k < NV; ++k) {
      double v = P.c[k];
      for (int j = 0; j < NV; ++j) v += P.Q[k][j] * x[j];
      rd[k] = v;
    }
    double mu = 0;
    for (int i = 0; i < m; ++i) {
      P.axpy(i, l[i], rd);
      rp[i] = P.dot(i, x) + s[i] - P.b[i];
      mu += s[i] * l[i];
    }
    const ConeEval ce = cone_eval(x, tv);
    rd[0] += ce.g0 * l[m];
    rd[1] += ce.g1 * l[m];
    if (tv >= 0) rd[tv] += ce.gt * l[m];
    rp[m] = ce.f + s[m];
    mu += s[m] * l[m];
    mu /= M;
    double rdn = 0, rpn = 0;
    for (int k = 0; k < NV; ++k) rdn = rmax(rdn, fabs(rd[k]));
    for (int i = 0; i < M; ++i) rpn = rmax(rpn, fabs(rp[i]));
#ifdef RDA_IPM_DEBUG
    printf("it %d rd %.2e rp %.2e mu %.2e x0 %g x1 %g\n", it, rdn, rpn, mu, x[0], x[1]);
#endif
    if (!(rdn == rdn) || !(mu == mu)) return false;
    acceptable = rdn < 1e-6 * scale && rpn < 1e-6 * scale && mu < 1e-7;
    if (rdn < 1e-9 * scale && rpn < 1e-9 * scale && mu < 1e-10) return true;
    if (mu < 1e-14) return acceptable;     // complementarity exhausted (rounding floor reached)
    // Newton matrix
    double H[NV][NV];
    for (int k = 0; k < NV; ++k)
      for (int j = 0; j < NV; ++j) H[k][j] = P.Q[k][j];
    for (int i = 0; i < m; ++i) P.rank1(i, l[i] / s[i], H);
    {
      double w = l[m] / s[m];
      H[0][0] += l[m] * ce.h00 + w * ce.g0 * ce.g0;
      H[1][0] += w * ce.g1 * ce.g0;
      H[1][1] += l[m] * ce.h11 + w * ce.g1 * ce.g1;
      if (tv >= 0) {   // tv > 1 always (lower triangle: row tv, columns 0, 1, tv)
        H[tv][0] += l[m] * ce.h0t + w * ce.gt * ce.g0;
        H[tv][1] += l[m] * ce.h1t + w * ce.gt * ce.g1;
        H[tv][tv] += l[m] * ce.htt + w * ce.gt * ce.gt;
      }
    }
    for (int k = 0; k < NV; ++k) H[k][k] += 1e-12;
    // affine right-hand side: -(rd + sum grad_i (l_i rp_i - rc_i)/s_i), rc_i = s_i l_i
    double ra[NV], rc[NV];
    for (int k = 0; k < NV; ++k) ra[k] = -rd[k];
    for (int i = 0; i < M; ++i) {
      double t = (l[i] * rp[i] - s[i] * l[i]) / s[i];
      if (i < m) {
        P.axpy(i, -t, ra);
      } else {
        ra[0] -= ce.g0 * t;
        ra[1] -= ce.g1 * t;
        if (tv >= 0) ra[tv] -= ce.gt * t;
      }
    }
    for (int k = 0; k < NV; ++k) rc[k] = ra[k];
    double Hc[NV][NV];
    for (int k = 0; k < NV; ++k)
      for (int j = 0; j < NV; ++j) Hc[k][j] = H[k][j];
    if (!chol_solve<NV>(Hc, ra, nullptr)) return acceptable;
    // affine step lengths
    double dsa[MC + 1], dla[MC + 1], aaff = 1.0;
    for (int i = 0; i < M; ++i) {
      double gd = 0;
      if (i < m) {
        gd = P.dot(i, ra);
      } else {
        gd = ce.g0 * ra[0] + ce.g1 * ra[1] + (tv >= 0 ? ce.gt * ra[tv] : 0.0);
      }
      dsa[i] = -rp[i] - gd;
      dla[i] = -(s[i] * l[i] + l[i] * dsa[i]) / s[i];
      if (dsa[i] < 0) aaff = rmin(aaff, -s[i] / dsa[i]);
      if (dla[i] < 0) aaff = rmin(aaff, -l[i] / dla[i]);
    }
    double mua = 0;
    for (int i = 0; i < M; ++i) mua += (s[i] + aaff * dsa[i]) * (l[i] + aaff * dla[i]);
    mua /= M;
    double sig = mua / mu;
    sig = sig * sig * sig;
    // corrector right-hand side
    for (int k = 0; k < NV; ++k) rc[k] = -rd[k];
    double rcs[MC + 1];
    for (int i = 0; i < M; ++i) {
      rcs[i] = s[i] * l[i] + dsa[i] * dla[i] - sig * mu;
      double t = (l[i] * rp[i] - rcs[i]) / s[i];
      if (i < m) {
        P.axpy(i, -t, rc);
      } else {
        rc[0] -= ce.g0 * t;
        rc[1] -= ce.g1 * t;
        if (tv >= 0) rc[tv] -= ce.gt * t;
      }
    }
    // Hc already holds the factor: only the triangular solves are needed
    for (int i = 0; i < NV; ++i) {
      double sv = rc[i];
      for (int k = 0; k < i; ++k) sv -= Hc[i][k] * rc[k];
      rc[i] = sv / Hc[i][i];
    }
    for (int i = NV - 1; i >= 0; --i) {
      double sv = rc[i];
      for (int k = i + 1; k < NV; ++k) sv -= Hc[k][i] * rc[k];
      rc[i] = sv / Hc[i][i];
    }
    double alpha = 1.0, ds[MC + 1], dl[MC + 1];
    for (int i = 0; i < M; ++i) {
      double gd = 0;
      if (i < m) {
        gd = P.dot(i, rc);
      } else {
        gd = ce.g0 * rc[0] + ce.g1 * rc[1] + (tv >= 0 ? ce.gt * rc[tv] : 0.0);
      }
      ds[i] = -rp[i] - gd;
      dl[i] = -(rcs[i] + l[i] * ds[i]) / s[i];
      if (ds[i] < 0) alpha = rmin(alpha, -0.995 * s[i] / ds[i]);
      if (dl[i] < 0) alpha = rmin(alpha, -0.995 * l[i] / dl[i]);
    }
    for (int k = 0; k < NV; ++k) x[k] += alpha * rc[k];
    for (int i = 0; i < M; ++i) {
      s[i] += alpha * ds[i];
      l[i] += alpha * dl[i];
    }
  }
  return acceptable;  // iteration cap reached
}

// ---------------------------------------------------------------------------------------------
// Feasible-start log-barrier method (damped Newton with backtracking) for the same problem class
// with the second-order-cone constraint |x01| <= x_tv (P.tv >= 0) in its self-concordant form
// -log(x_tv^2 - |x01|^2).  Slower than tiny_ipm but globally convergent; used for disc obstacles.
// ---------------------------------------------------------------------------------------------
template <int NV, int MC>
RDA_HD double barrier_value(const TinyQP<NV, MC>& P, const double* x, double t) {
  double f = 0;
  for (int k = 0; k < NV; ++k) {
    double qx = 0;
    for (int j = 0; j < NV; ++j) qx += P.Q[k][j] * x[j];
    f += x[k] * (0.5 * qx + P.c[k]);
  }
  f *= t;
  for (int i = 0; i < P.m; ++i) {
    double sl = P.b[i] - P.dot(i, x);
    if (!(sl > 0)) return 1e300;
    f -= log(sl);
  }
  double tq = P.tv >= 0 ? x[P.tv] : 1.0;
  double psi = tq * tq - x[0] * x[0] - x[1] * x[1];
  if (!(psi > 0) || !(tq > 0)) return 1e300;
  return f - log(psi);
}

template <int NV, int MC>
RDA_HD_NOINLINE bool tiny_barrier(const TinyQP<NV, MC>& P, double* x /* strictly feasible */) {
  const int m = P.m, tv = P.tv;
  double t = 1.0;
  for (int outer = 0; outer < 14; ++outer, t *= 8.0) {
    for (int it = 0; it < 30; ++it) {
      double g[NV], H[NV][NV];
      for (int k = 0; k < NV; ++k) {
        double qx = P.c[k];
        for (int j = 0; j < NV; ++j) { qx += P.Q[k][j] * x[j]; H[k][j] = t * P.Q[k][j]; }
        g[k] = t * qx;
      }
      for (int i = 0; i < m; ++i) {
        double inv = 1.0 / (P.b[i] - P.dot(i, x));
        P.axpy(i, inv, g);
        P.rank1(i, inv * inv, H);
      }
      {
        const double tq = tv >= 0 ? x[tv] : 1.0;
        double psi = tq * tq - x[0] * x[0] - x[1] * x[1];
        double gp[3] = {-2 * x[0], -2 * x[1], 2 * tq};          // grad psi on (0, 1, tv)
        const int id[3] = {0, 1, tv};
        const int na = tv >= 0 ? 3 : 2;
        double ip = 1.0 / psi;
        for (int a = 0; a < na; ++a) {
          g[id[a]] -= gp[a] * ip;
          for (int b2 = 0; b2 <= a; ++b2) H[id[a]][id[b2]] += gp[a] * gp[b2] * ip * ip;
        }
        H[0][0] += 2 * ip; H[1][1] += 2 * ip;
        if (tv >= 0) H[tv][tv] -= 2 * ip;
      }
      for (int k = 0; k < NV; ++k) H[k][k] += 1e-13 * (1.0 + H[k][k]);
      double dx[NV];
      for (int k = 0; k < NV; ++k) dx[k] = -g[k];
      if (!chol_solve<NV>(H, dx, nullptr)) return false;
      double lam2 = 0;
      for (int k = 0; k < NV; ++k) lam2 -= g[k] * dx[k];
      if (!(lam2 == lam2)) return false;
      if (lam2 < 1e-9) break;
      double f0 = barrier_value<NV, MC>(P, x, t), step = 1.0;
      double xn[NV];
      bool moved = false;
      for (int bt = 0; bt < 50; ++bt, step *= 0.5) {
        for (int k = 0; k < NV; ++k) xn[k] = x[k] + step * dx[k];
        double f1 = barrier_value<NV, MC>(P, xn, t);
        if (f1 <= f0 - 0.1 * step * lam2) { moved = true; break; }
      }
      if (!moved) break;
      for (int k = 0; k < NV; ++k) x[k] = xn[k];
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// Geometry of one cell, relative to the robot reference point.
// ---------------------------------------------------------------------------------------------
template <typename Real>
struct CellGeom {
  int kind, ne;
  Real vx[RDA_MAX_EDGE], vy[RDA_MAX_EDGE];     // polygon vertices (vertex i joins rows i-1, i)
  Real nx[RDA_MAX_EDGE], ny[RDA_MAX_EDGE];     // unit outward normals of rows
  Real inv_norm[RDA_MAX_EDGE];                 // 1/|A_i|
  Real cx, cy, rad;                            // disc
  Real yx[RDA_MAX_ROBOT_EDGE], yy[RDA_MAX_ROBOT_EDGE];  // robot vertices rotated into the world frame
  Real mx[RDA_MAX_ROBOT_EDGE], my[RDA_MAX_ROBOT_EDGE];  // robot edge normals in the world frame
};

template <typename Real>
RDA_HD Real support_obs(const CellGeom<Real>& g, Real vx, Real vy, int* arg) {
  if (g.kind == RDA_OBS_CIRCLE) {
    *arg = 0;
    return vx * g.cx + vy * g.cy + g.rad * sqrt_(vx * vx + vy * vy);
  }
  Real best = -1e30f;
  int bi = 0;
  for (int i = 0; i < g.ne; ++i) {
    Real s = vx * g.vx[i] + vy * g.vy[i];
    if (s > best) { best = s; bi = i; }
  }
  *arg = bi;
  return best;
}

template <typename Real>
RDA_HD Real support_rob(const RobotGeom& rb, Real gx, Real gy, int* arg) {
  Real best = -1e30f;
  int bj = 0;
  for (int j = 0; j < rb.R; ++j) {
    Real s = gx * (Real)rb.yx[j] + gy * (Real)rb.yy[j];
    if (s > best) { best = s; bj = j; }
  }
  *arg = bj;
  return best;
}

// g (body frame) inside the normal cone of body vertex j?
template <typename Real>
RDA_HD bool in_cone_rob(const RobotGeom& rb, int j, Real gx, Real gy, Real tol) {
  const int R = rb.R;
  int jp = (j + R - 1) % R, jn = (j + 1) % R;
  Real epx = (Real)rb.yx[j] - (Real)rb.yx[jp], epy = (Real)rb.yy[j] - (Real)rb.yy[jp];
  Real enx = (Real)rb.yx[jn] - (Real)rb.yx[j], eny = (Real)rb.yy[jn] - (Real)rb.yy[j];
  Real gn = sqrt_(gx * gx + gy * gy);
  Real a = gx * epx + gy * epy, b = gx * enx + gy * eny;
  return a >= -tol * gn * sqrt_(epx * epx + epy * epy) && b <= tol * gn * sqrt_(enx * enx + eny * eny);
}

// FAST_ONLY = true compiles the closed-form paths only and reports CELL_NEEDS_SLOW (no outputs
// written) for cells that need the interior point method (first pass of the two-pass kernel).
template <typename Real, bool FAST_ONLY = false>
RDA_HD void cell_solve(const RobotGeom& rb, int kind, int E, const float* A, const float* b,
                       Real px, Real py, Real cphi, Real sphi, Real dbar, Real zeta, Real xi0,
                       Real xi1, Real ro2, Real theta, CellOut<Real>& out) {
  const int R = rb.R;
  const Real k0 = dbar - zeta;
  const Real eps = sizeof(Real) == 4 ? (Real)1e-5 : (Real)1e-11;
  CellGeom<Real> g;
  g.kind = kind;
  for (int j = 0; j < R; ++j) {
    Real yx = rb.yx[j], yy = rb.yy[j];
    g.yx[j] = cphi * yx - sphi * yy;
    g.yy[j] = sphi * yx + cphi * yy;
    Real nx = rb.nx[j], ny = rb.ny[j];
    g.mx[j] = cphi * nx - sphi * ny;
    g.my[j] = sphi * nx + cphi * ny;
  }
  // ---- obstacle in coordinates relative to p ------------------------------------------------
  int ne = 0;
  Real brel[RDA_MAX_EDGE];
  if (kind == RDA_OBS_CIRCLE) {
    g.cx = (Real)b[0] - px;
    g.cy = (Real)b[1] - py;
    g.rad = -(Real)b[2];
    g.ne = 0;
  } else {
    for (int i = 0; i < E; ++i) {
      Real ax = A[2 * i], ay = A[2 * i + 1];
      Real n2 = ax * ax + ay * ay;
      if (!(n2 > 0)) break;
      Real inv = rsqrt_(n2);
      g.nx[i] = ax * inv;
      g.ny[i] = ay * inv;
      g.inv_norm[i] = inv;
      brel[i] = ((Real)b[i] - ax * px - ay * py) * inv;
      ne = i + 1;
    }
    g.ne = ne;
    for (int i = 0; i < ne; ++i) {
      int a = (i + ne - 1) % ne;
      Real det = g.nx[a] * g.ny[i] - g.ny[a] * g.nx[i];
      Real inv = (Real)1 / det;
      g.vx[i] = (brel[a] * g.ny[i] - brel[i] * g.ny[a]) * inv;
      g.vy[i] = (g.nx[a] * brel[i] - g.nx[i] * brel[a]) * inv;
    }
  }
  // ---- closest pair / separation -------------------------------------------------------------
  bool sep = false;
  Real best = 1e30f, bdx = 0, bdy = 0;
  Real byx = 0, byy = 0;            // robot-side point of the closest pair, body frame
  Real rob_in[RDA_MAX_ROBOT_EDGE], obs_in[RDA_MAX_EDGE];
  bool have_in = false;
  Real dj2[RDA_MAX_ROBOT_EDGE], djx[RDA_MAX_ROBOT_EDGE], djy[RDA_MAX_ROBOT_EDGE];
  for (int j = 0; j < R; ++j) dj2[j] = 1e30f;
  if (kind == RDA_OBS_CIRCLE) {
    bool inside = true;
    for (int j = 0; j < R; ++j) {
      int jn = (j + 1) % R;
      Real fx = g.yx[jn] - g.yx[j], fy = g.yy[jn] - g.yy[j];
      Real rx = g.cx - g.yx[j], ry = g.cy - g.yy[j];
      if (g.mx[j] * rx + g.my[j] * ry > 0) inside = false;
      Real t = rclamp((rx * fx + ry * fy) / (fx * fx + fy * fy), (Real)0, (Real)1);
      Real dx = -(rx - t * fx), dy = -(ry - t * fy);       // robot point minus centre
      Real d2 = dx * dx + dy * dy;
      if (d2 < best) {
        best = d2; bdx = dx; bdy = dy;
        byx = (Real)rb.yx[j] + t * ((Real)rb.yx[jn] - (Real)rb.yx[j]);
        byy = (Real)rb.yy[j] + t * ((Real)rb.yy[jn] - (Real)rb.yy[j]);
      }
      Real vx_ = g.yx[j] - g.cx, vy_ = g.yy[j] - g.cy;     // robot vertex minus centre
      Real dv = sqrt_(vx_ * vx_ + vy_ * vy_);
      Real dd = dv - g.rad;
      if (dv > eps && dd > 0) {
        dj2[j] = dd * dd;
        djx[j] = vx_ / dv * dd;
        djy[j] = vy_ / dv * dd;
      } else {
        dj2[j] = 0; djx[j] = 0; djy[j] = 0;
      }
    }
    Real dc = sqrt_(best);
    sep = (!inside) && (dc > g.rad + eps);
    if (sep) {
      Real dd = dc - g.rad;
      bdx = bdx / dc * dd;
      bdy = bdy / dc * dd;
      best = dd * dd;
    }
  } else {
    for (int j = 0; j < R; ++j) rob_in[j] = -1e30f;
    for (int i = 0; i < ne; ++i) {
      int in = (i + 1) % ne;
      Real ex = g.vx[in] - g.vx[i], ey = g.vy[in] - g.vy[i];
      Real ie2 = (Real)1 / (ex * ex + ey * ey);
      Real mins = 1e30f;
      for (int j = 0; j < R; ++j) {
        Real rx = g.yx[j] - g.vx[i], ry = g.yy[j] - g.vy[i];
        Real sd = g.nx[i] * rx + g.ny[i] * ry;
        mins = rmin(mins, sd);
        rob_in[j] = rmax(rob_in[j], sd);        // <= 0 for every edge: robot vertex j inside O
        Real t = rclamp((rx * ex + ry * ey) * ie2, (Real)0, (Real)1);
        Real dx = rx - t * ex, dy = ry - t * ey;
        Real d2 = dx * dx + dy * dy;
        if (d2 < dj2[j]) { dj2[j] = d2; djx[j] = dx; djy[j] = dy; }
      }
      if (mins > eps) sep = true;
    }
    for (int j = 0; j < R; ++j)
      if (dj2[j] < best) { best = dj2[j]; bdx = djx[j]; bdy = djy[j]; byx = rb.yx[j]; byy = rb.yy[j]; }
    for (int j = 0; j < R; ++j) {
      int jn = (j + 1) % R;
      Real fx = g.yx[jn] - g.yx[j], fy = g.yy[jn] - g.yy[j];
      Real if2 = (Real)1 / (fx * fx + fy * fy);
      Real mins = 1e30f;
      for (int i = 0; i < ne; ++i) {
        Real rx = g.vx[i] - g.yx[j], ry = g.vy[i] - g.yy[j];
        Real sd = g.mx[j] * rx + g.my[j] * ry;
        mins = rmin(mins, sd);
        obs_in[i] = (j == 0) ? sd : rmax(obs_in[i], sd);   // <= 0: obstacle vertex i inside the robot
        Real t = rclamp((rx * fx + ry * fy) * if2, (Real)0, (Real)1);
        Real dx = -(rx - t * fx), dy = -(ry - t * fy);
        Real d2 = dx * dx + dy * dy;
        if (d2 < best) {
          best = d2; bdx = dx; bdy = dy;
          byx = (Real)rb.yx[j] + t * ((Real)rb.yx[jn] - (Real)rb.yx[j]);
          byy = (Real)rb.yy[j] + t * ((Real)rb.yy[jn] - (Real)rb.yy[j]);
        }
      }
      if (mins > eps) sep = true;
    }
  }
  // ---- candidate (v, g) -------------------------------------------------------------------------
  Real v0 = 0, v1 = 0, g0 = 0, g1 = 0;
  bool exact_zero_q = false, have = false;
  int path = CELL_FAILED;
  const bool xi_zero = (xi0 == (Real)0) && (xi1 == (Real)0);
  if (sep && xi_zero) {
    Real dist = sqrt_(best);
    if (dist - k0 >= 0) {
      v0 = bdx / dist; v1 = bdy / dist;
      g0 = -(cphi * v0 + sphi * v1);
      g1 = -(-sphi * v0 + cphi * v1);
      exact_zero_q = true; have = true; path = CELL_FAST_INACTIVE;
    }
  }
  if (!have && !sep && xi_zero && k0 <= 0) {
    // overlapping sets, no tilt: max margin is 0 at v = 0 (stuff = -k0 >= 0)
    exact_zero_q = true; have = true; path = CELL_OVERLAP_FREE;
  }
  if (!have && sep) {
    const Real tolc = sizeof(Real) == 4 ? (Real)1e-5 : (Real)1e-11;
    for (int j = 0; j < R && !have; ++j) {
      Real dj = sqrt_(dj2[j]);
      if (!(dj > eps)) continue;
      Real vjx = djx[j] / dj, vjy = djy[j] / dj;
      Real yx = rb.yx[j], yy = rb.yy[j];
      Real Dj = dj + xi0 * yx + xi1 * yy - k0;
      Real rvx = cphi * vjx + sphi * vjy, rvy = -sphi * vjx + cphi * vjy;   // R'v
      if (Dj >= 0) {
        Real cgx = -rvx - xi0, cgy = -rvy - xi1;
        if (in_cone_rob<Real>(rb, j, cgx, cgy, tolc)) {
          v0 = vjx; v1 = vjy; g0 = cgx; g1 = cgy;
          exact_zero_q = true; have = true; path = CELL_FAST_VERTEX;
        }
      } else {
        Real tau = -Dj / ((Real)1 + (yx * yx + yy * yy) / ro2);
        Real qx = -tau * yx / ro2, qy = -tau * yy / ro2;
        Real cgx = qx - rvx - xi0, cgy = qy - rvy - xi1;
        if (in_cone_rob<Real>(rb, j, cgx, cgy, tolc)) {
          v0 = vjx; v1 = vjy; g0 = cgx; g1 = cgy;
          have = true; path = CELL_FAST_VERTEX;
        }
      }
    }
  }
#ifdef RDA_CELL_STATS
  if (!have) {
    extern long long g_cell_stats[8];
    __sync_fetch_and_add(&g_cell_stats[(sep ? 0 : 2) + (xi_zero ? 0 : 1)], 1);
  }
#endif
  if (FAST_ONLY) {
    if (!have) { out.path = CELL_NEEDS_SLOW; return; }
  } else if (!have) {
    // ---- slow path: interior point in float64 ---------------------------------------------------
    // Polygon obstacle: sigma_O(v) = max_i v.x_i, |v| <= 1 (ball constraint).
    // Disc obstacle:    sigma_O(v) = v.c + rad*tv with |v| <= tv <= 1 (cone constraint, extra
    //                   variable tv), exact also when the robot overlaps the disc.
    const bool circ = (kind == RDA_OBS_CIRCLE);
    const int nv_o = circ ? 1 : ne;
    double ox[RDA_MAX_EDGE], oy[RDA_MAX_EDGE];
    const double k0d = (double)k0, radd = circ ? (double)g.rad : 0.0;
    if (circ) { ox[0] = g.cx; oy[0] = g.cy; }
    else for (int i = 0; i < ne; ++i) { ox[i] = g.vx[i]; oy[i] = g.vy[i]; }
    const double c_ = cphi, s_ = sphi, x0 = xi0, x1 = xi1;
    bool ok = true, inactive = false;
    double va = 0, vb = 0, ga = 0, gb = 0;
    // The max margin c* = min_{x in O, y in Rob} |P(y) - x| + xi.y - k0 is bounded above by its value
    // at the closest pair (disjoint sets): if that is negative the hinge is active for sure and
    // stage A can be skipped.
    bool need_a = true;
    if (sep) {
      double ub = sqrt((double)best) + x0 * (double)byx + x1 * (double)byy - k0d;
      if (ub < -1e-9) need_a = false;
    } else if (!circ) {
      // overlapping polygons: any common point y gives the bound xi.y - k0
      double ub = 1e300;
      for (int j = 0; j < R; ++j)
        if (rob_in[j] <= 0) ub = rmin(ub, x0 * (double)rb.yx[j] + x1 * (double)rb.yy[j]);
      for (int i = 0; i < ne; ++i)
        if (obs_in[i] <= 0) {
          double wx = g.vx[i], wy = g.vy[i];
          ub = rmin(ub, x0 * (c_ * wx + s_ * wy) + x1 * (-s_ * wx + c_ * wy));
        }
      if (ub - k0d < -1e-9) need_a = false;
    }
    (void)have_in;
#ifdef RDA_CELL_STATS
    if (need_a) { extern long long g_cell_stats[8]; __sync_fetch_and_add(&g_cell_stats[6], 1); }
#endif
    if (need_a) {  // stage A: max margin with Hm + xi = 0; x = (v0, v1, so, sr, tv)
      constexpr int NVA = 5;
      TinyQP<NVA, 2 * RDA_MAX_EDGE + 2> P;
      for (int k = 0; k < NVA; ++k) { for (int j = 0; j < NVA; ++j) P.Q[k][j] = 0; P.c[k] = 0; }
      P.c[2] = 1; P.c[3] = 1;
      P.m = 0;
      for (int i = 0; i < nv_o; ++i) P.row(4, 0, ox[i], 1, oy[i], 2, -1.0, 4, radd, 0.0);   // v.x_i + rad tv <= so
      for (int j = 0; j < R; ++j) {
        // g.y_j <= sr with g = -R'v - xi :  -(R y_j).v - sr <= xi.y_j
        double yx = rb.yx[j], yy = rb.yy[j];
        P.row(3, 0, -(c_ * yx - s_ * yy), 1, -(s_ * yx + c_ * yy), 3, -1.0, 0, 0.0, x0 * yx + x1 * yy);
      }
      P.row(1, 4, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, 1.0);          // tv <= 1
      P.row(1, 4, -1.0, 0, 0.0, 0, 0.0, 0, 0.0, 0.0);         // tv >= 0
      P.tv = circ ? 4 : -1;
      double hmax = 0;
      for (int j = 0; j < R; ++j) hmax = rmax(hmax, fabs(x0 * (double)rb.yx[j] + x1 * (double)rb.yy[j]));
      double xs[NVA] = {0, 0, 1.0 + radd, 1.0 + hmax, 0.5};
      ok = !circ && tiny_ipm<NVA, 2 * RDA_MAX_EDGE + 2>(P, xs);
      if (!ok) {   // discs, and the rare polygon cell on which the primal-dual iteration cycles
        xs[0] = 0; xs[1] = 0; xs[2] = 1.0 + radd; xs[3] = 1.0 + hmax; xs[4] = 0.5;
        ok = tiny_barrier<NVA, 2 * RDA_MAX_EDGE + 2>(P, xs);
      }
      double cst = -xs[2] - xs[3] - k0d;
      if (ok && cst >= 0) {
        inactive = true;
        va = xs[0]; vb = xs[1];
        ga = -(c_ * va + s_ * vb) - x0;
        gb = -(-s_ * va + c_ * vb) - x1;
        path = CELL_SLOW_A;
#ifdef RDA_CELL_STATS
        { extern long long g_cell_stats[8]; __sync_fetch_and_add(&g_cell_stats[4 + (sep ? 0 : 1)], 1); }
#endif
      }
    }
    if (ok && !inactive) {  // stage B: x = (v0, v1, g0, g1, so, sr, w, tv)
      constexpr int NVB = 8;
      TinyQP<NVB, 2 * RDA_MAX_EDGE + 3> P;
      for (int k = 0; k < NVB; ++k) { for (int j = 0; j < NVB; ++j) P.Q[k][j] = 0; P.c[k] = 0; }
      // ro2/2 |g + R'v + xi|^2 : with u = (v, g), q = M u + xi, M = [R' I]
      const double r2 = ro2;
      const double Mx[4] = {c_, s_, 1, 0}, My[4] = {-s_, c_, 0, 1};
      for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 4; ++j) P.Q[k][j] = r2 * (Mx[k] * Mx[j] + My[k] * My[j]);
      for (int k = 0; k < 4; ++k) P.c[k] = r2 * (Mx[k] * x0 + My[k] * x1);
      P.Q[6][6] = 1.0;   // 1/2 w^2  (ro1 == 1 inside LamMuZ, rda_solver.py:257)
      P.m = 0;
      for (int i = 0; i < nv_o; ++i) P.row(4, 0, ox[i], 1, oy[i], 4, -1.0, 7, radd, 0.0);
      for (int j = 0; j < R; ++j) P.row(3, 2, (double)rb.yx[j], 3, (double)rb.yy[j], 5, -1.0, 0, 0.0, 0.0);
      P.row(3, 4, 1.0, 5, 1.0, 6, -1.0, 0, 0.0, -k0d);        // so + sr + k0 <= w
      P.row(1, 7, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, 1.0);          // tv <= 1
      P.row(1, 7, -1.0, 0, 0.0, 0, 0.0, 0, 0.0, 0.0);         // tv >= 0
      P.tv = circ ? 7 : -1;
      double so0 = 1.0 + radd;
      const double w0 = rmax(so0 + 2.0 + k0d, 1.0);
      double xs[NVB] = {0, 0, 0, 0, so0, 1.0, w0, 0.5};
      ok = !circ && tiny_ipm<NVB, 2 * RDA_MAX_EDGE + 3>(P, xs);
      if (!ok) {
        xs[0] = xs[1] = xs[2] = xs[3] = 0; xs[4] = so0; xs[5] = 1.0; xs[6] = w0; xs[7] = 0.5;
        ok = tiny_barrier<NVB, 2 * RDA_MAX_EDGE + 3>(P, xs);
      }
      va = xs[0]; vb = xs[1]; ga = xs[2]; gb = xs[3];
      path = CELL_SLOW_B;
    }
    if (!ok) {
      path = CELL_FAILED;
    } else {
      v0 = (Real)va; v1 = (Real)vb; g0 = (Real)ga; g1 = (Real)gb;
      exact_zero_q = inactive;
      have = true;
    }
  }
  // ---- epilogue: multipliers, updates, su-QP inputs ------------------------------------------
  for (int i = 0; i < RDA_MAX_EDGE; ++i) out.lam[i] = 0;
  for (int j = 0; j < RDA_MAX_ROBOT_EDGE; ++j) out.mu[j] = 0;
  out.path = path;
  if (!have) {
    // keep-previous-iterate rule (rda_solver.py:791-793) is applied by the caller
    out.z = 0; out.zeta_new = zeta; out.xi0_new = xi0; out.xi1_new = xi1;
    out.ax = out.ay = out.c0 = out.gx = out.gy = out.hm0 = out.hm1 = 0;
    return;
  }
  int io = 0, jr = 0;
  Real sO = support_obs<Real>(g, v0, v1, &io);
  Real sR = support_rob<Real>(rb, g0, g1, &jr);
  Real vn = sqrt_(v0 * v0 + v1 * v1);
  if (kind == RDA_OBS_CIRCLE) {
    out.lam[0] = v0; out.lam[1] = v1; out.lam[2] = -vn;          // (v, -|v|), mpc.py:440-458
  } else if (vn > 0) {
    int a = (io + ne - 1) % ne, bb = io;
    Real det = g.nx[a] * g.ny[bb] - g.ny[a] * g.nx[bb];
    Real al = (v0 * g.ny[bb] - v1 * g.nx[bb]) / det;
    Real be = (g.nx[a] * v1 - g.ny[a] * v0) / det;
    out.lam[a] = rmax(al, (Real)0) * g.inv_norm[a];
    out.lam[bb] = rmax(be, (Real)0) * g.inv_norm[bb];
  }
  Real gn = sqrt_(g0 * g0 + g1 * g1);
  if (gn > 0) {
    int a = (jr + R - 1) % R, bb = jr;
    Real nax = rb.nx[a], nay = rb.ny[a], nbx = rb.nx[bb], nby = rb.ny[bb];
    Real det = nax * nby - nay * nbx;
    Real al = (g0 * nby - g1 * nbx) / det;
    Real be = (nax * g1 - nay * g0) / det;
    out.mu[a] = rmax(al, (Real)0) / (Real)rb.gnorm[a];
    out.mu[bb] = rmax(be, (Real)0) / (Real)rb.gnorm[bb];
  }
  Real marg = -sO - sR;                 // lam'(A p - b) - mu'h
  Real stuff = marg - k0;
  Real z = theta * rmax(stuff, (Real)0);
  Real q0, q1;
  if (exact_zero_q) { q0 = 0; q1 = 0; }
  else {
    q0 = g0 + (cphi * v0 + sphi * v1) + xi0;
    q1 = g1 + (-sphi * v0 + cphi * v1) + xi1;
  }
  out.z = z;
  out.zeta_new = stuff - z;             // zeta + Im - d - z  (:666)
  out.xi0_new = q0;                     // xi + Hm            (:683)
  out.xi1_new = q1;
  out.hm0 = q0 - xi0;
  out.hm1 = q1 - xi1;
  out.ax = v0;
  out.ay = v1;
  out.c0 = marg - z + out.zeta_new;     // a.p - lam'b - mu'h - z + zeta   (Im_su without -d, :846-851)
  out.gx = g0 + q0;                     // mu'G + xi                       (:868)
  out.gy = g1 + q1;
}

}  // namespace rda
