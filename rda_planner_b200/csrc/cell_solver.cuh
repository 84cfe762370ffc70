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
       CELL_OVERLAP_FREE = 4, CELL_FAILED = 5 };

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
  double a[MC][NV];
  double b[MC];
  int m;
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

template <int NV, int MC>
RDA_HD_NOINLINE bool tiny_ipm(const TinyQP<NV, MC>& P, double* x /* in: strictly feasible start */) {
  const int m = P.m;
  double s[MC + 1], l[MC + 1];
  for (int i = 0; i < m; ++i) {
    double ax = 0;
    for (int k = 0; k < NV; ++k) ax += P.a[i][k] * x[k];
    s[i] = rmax(P.b[i] - ax, 1e-3);
    l[i] = 1.0 / s[i];
  }
  s[m] = rmax(1.0 - x[0] * x[0] - x[1] * x[1], 1e-3);
  l[m] = 1.0 / s[m];
  const int M = m + 1;
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
      double ax = 0;
      for (int k = 0; k < NV; ++k) {
        ax += P.a[i][k] * x[k];
        rd[k] += P.a[i][k] * l[i];
      }
      rp[i] = ax + s[i] - P.b[i];
      mu += s[i] * l[i];
    }
    rd[0] += 2 * x[0] * l[m];
    rd[1] += 2 * x[1] * l[m];
    rp[m] = x[0] * x[0] + x[1] * x[1] - 1.0 + s[m];
    mu += s[m] * l[m];
    mu /= M;
    double rdn = 0, rpn = 0;
    for (int k = 0; k < NV; ++k) rdn = rmax(rdn, fabs(rd[k]));
    for (int i = 0; i < M; ++i) rpn = rmax(rpn, fabs(rp[i]));
    if (rdn < 1e-10 && rpn < 1e-10 && mu < 1e-11) return true;
    if (!(rdn == rdn) || !(mu == mu)) return false;
    // Newton matrix
    double H[NV][NV];
    for (int k = 0; k < NV; ++k)
      for (int j = 0; j < NV; ++j) H[k][j] = P.Q[k][j];
    for (int i = 0; i < m; ++i) {
      double w = l[i] / s[i];
      for (int k = 0; k < NV; ++k) {
        double wk = w * P.a[i][k];
        if (wk != 0)
          for (int j = 0; j <= k; ++j) H[k][j] += wk * P.a[i][j];
      }
    }
    {
      double w = l[m] / s[m], g0 = 2 * x[0], g1 = 2 * x[1];
      H[0][0] += 2 * l[m] + w * g0 * g0;
      H[1][0] += w * g1 * g0;
      H[1][1] += 2 * l[m] + w * g1 * g1;
    }
    for (int k = 0; k < NV; ++k) H[k][k] += 1e-12;
    // affine right-hand side: -(rd + sum grad_i (l_i rp_i - rc_i)/s_i), rc_i = s_i l_i
    double ra[NV], rc[NV];
    for (int k = 0; k < NV; ++k) ra[k] = -rd[k];
    for (int i = 0; i < M; ++i) {
      double t = (l[i] * rp[i] - s[i] * l[i]) / s[i];
      if (i < m) {
        for (int k = 0; k < NV; ++k) ra[k] -= P.a[i][k] * t;
      } else {
        ra[0] -= 2 * x[0] * t;
        ra[1] -= 2 * x[1] * t;
      }
    }
    for (int k = 0; k < NV; ++k) rc[k] = ra[k];
    double Hc[NV][NV];
    for (int k = 0; k < NV; ++k)
      for (int j = 0; j < NV; ++j) Hc[k][j] = H[k][j];
    if (!chol_solve<NV>(Hc, ra, nullptr)) return false;
    // affine step lengths
    double dsa[MC + 1], dla[MC + 1], aaff = 1.0;
    for (int i = 0; i < M; ++i) {
      double gd = 0;
      if (i < m) {
        for (int k = 0; k < NV; ++k) gd += P.a[i][k] * ra[k];
      } else {
        gd = 2 * x[0] * ra[0] + 2 * x[1] * ra[1];
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
        for (int k = 0; k < NV; ++k) rc[k] -= P.a[i][k] * t;
      } else {
        rc[0] -= 2 * x[0] * t;
        rc[1] -= 2 * x[1] * t;
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
        for (int k = 0; k < NV; ++k) gd += P.a[i][k] * rc[k];
      } else {
        gd = 2 * x[0] * rc[0] + 2 * x[1] * rc[1];
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
  return true;  // iteration cap reached: the iterate is still the best available point
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

template <typename Real>
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
      if (d2 < best) { best = d2; bdx = dx; bdy = dy; }
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
    for (int i = 0; i < ne; ++i) {
      int in = (i + 1) % ne;
      Real ex = g.vx[in] - g.vx[i], ey = g.vy[in] - g.vy[i];
      Real ie2 = (Real)1 / (ex * ex + ey * ey);
      Real mins = 1e30f;
      for (int j = 0; j < R; ++j) {
        Real rx = g.yx[j] - g.vx[i], ry = g.yy[j] - g.vy[i];
        Real sd = g.nx[i] * rx + g.ny[i] * ry;
        mins = rmin(mins, sd);
        Real t = rclamp((rx * ex + ry * ey) * ie2, (Real)0, (Real)1);
        Real dx = rx - t * ex, dy = ry - t * ey;
        Real d2 = dx * dx + dy * dy;
        if (d2 < dj2[j]) { dj2[j] = d2; djx[j] = dx; djy[j] = dy; }
      }
      if (mins > eps) sep = true;
    }
    for (int j = 0; j < R; ++j)
      if (dj2[j] < best) { best = dj2[j]; bdx = djx[j]; bdy = djy[j]; }
    for (int j = 0; j < R; ++j) {
      int jn = (j + 1) % R;
      Real fx = g.yx[jn] - g.yx[j], fy = g.yy[jn] - g.yy[j];
      Real if2 = (Real)1 / (fx * fx + fy * fy);
      Real mins = 1e30f;
      for (int i = 0; i < ne; ++i) {
        Real rx = g.vx[i] - g.yx[j], ry = g.vy[i] - g.yy[j];
        Real sd = g.mx[j] * rx + g.my[j] * ry;
        mins = rmin(mins, sd);
        Real t = rclamp((rx * fx + ry * fy) * if2, (Real)0, (Real)1);
        Real dx = -(rx - t * fx), dy = -(ry - t * fy);
        Real d2 = dx * dx + dy * dy;
        if (d2 < best) { best = d2; bdx = dx; bdy = dy; }
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
  if (!have) {
    // ---- slow path: interior point in float64 ---------------------------------------------------
    // A disc obstacle is handled as its centre point with k0 + rad (exact whenever the optimal
    // contact point lies outside the disc).
    const int nv_o = (kind == RDA_OBS_CIRCLE) ? 1 : ne;
    double ox[RDA_MAX_EDGE], oy[RDA_MAX_EDGE];
    double k0d = (double)k0;
    if (kind == RDA_OBS_CIRCLE) { ox[0] = g.cx; oy[0] = g.cy; k0d += (double)g.rad; }
    else for (int i = 0; i < ne; ++i) { ox[i] = g.vx[i]; oy[i] = g.vy[i]; }
    const double c_ = cphi, s_ = sphi, x0 = xi0, x1 = xi1;
    bool ok = true, inactive = false;
    double va = 0, vb = 0, ga = 0, gb = 0;
    {  // stage A: max margin with Hm + xi = 0; x = (v0, v1, so, sr)
      TinyQP<4, 2 * RDA_MAX_EDGE> P;
      for (int k = 0; k < 4; ++k) { for (int j = 0; j < 4; ++j) P.Q[k][j] = 0; P.c[k] = 0; }
      P.c[2] = 1; P.c[3] = 1;
      int m = 0;
      for (int i = 0; i < nv_o; ++i, ++m) {
        P.a[m][0] = ox[i]; P.a[m][1] = oy[i]; P.a[m][2] = -1; P.a[m][3] = 0; P.b[m] = 0;
      }
      for (int j = 0; j < R; ++j, ++m) {
        // g.y_j <= sr with g = -R'v - xi :  -(R y_j).v - sr <= xi.y_j
        double yx = rb.yx[j], yy = rb.yy[j];
        P.a[m][0] = -(c_ * yx - s_ * yy); P.a[m][1] = -(s_ * yx + c_ * yy);
        P.a[m][2] = 0; P.a[m][3] = -1; P.b[m] = x0 * yx + x1 * yy;
      }
      P.m = m;
      double hmax = 0;
      for (int j = 0; j < R; ++j) hmax = rmax(hmax, fabs(x0 * (double)rb.yx[j] + x1 * (double)rb.yy[j]));
      double xs[4] = {0, 0, 1.0, 1.0 + hmax};
      ok = tiny_ipm<4, 2 * RDA_MAX_EDGE>(P, xs);
      double cst = -xs[2] - xs[3] - k0d;
      if (ok && cst >= 0) {
        inactive = true;
        va = xs[0]; vb = xs[1];
        ga = -(c_ * va + s_ * vb) - x0;
        gb = -(-s_ * va + c_ * vb) - x1;
        path = CELL_SLOW_A;
      }
    }
    if (ok && !inactive) {  // stage B: x = (v0, v1, g0, g1, so, sr, w)
      TinyQP<7, 2 * RDA_MAX_EDGE + 1> P;
      for (int k = 0; k < 7; ++k) { for (int j = 0; j < 7; ++j) P.Q[k][j] = 0; P.c[k] = 0; }
      // ro2/2 |g + R'v + xi|^2 : with u = (v, g), q = M u + xi, M = [R' I]
      const double r2 = ro2;
      const double Mx[4] = {c_, s_, 1, 0}, My[4] = {-s_, c_, 0, 1};
      for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 4; ++j) P.Q[k][j] = r2 * (Mx[k] * Mx[j] + My[k] * My[j]);
      for (int k = 0; k < 4; ++k) P.c[k] = r2 * (Mx[k] * x0 + My[k] * x1);
      P.Q[6][6] = 1.0;   // 1/2 w^2  (ro1 == 1 inside LamMuZ, rda_solver.py:257)
      int m = 0;
      for (int i = 0; i < nv_o; ++i, ++m) {
        for (int k = 0; k < 7; ++k) P.a[m][k] = 0;
        P.a[m][0] = ox[i]; P.a[m][1] = oy[i]; P.a[m][4] = -1; P.b[m] = 0;
      }
      for (int j = 0; j < R; ++j, ++m) {
        for (int k = 0; k < 7; ++k) P.a[m][k] = 0;
        P.a[m][2] = rb.yx[j]; P.a[m][3] = rb.yy[j]; P.a[m][5] = -1; P.b[m] = 0;
      }
      for (int k = 0; k < 7; ++k) P.a[m][k] = 0;
      P.a[m][4] = 1; P.a[m][5] = 1; P.a[m][6] = -1; P.b[m] = -k0d;   // so + sr + k0 <= w
      ++m;
      P.m = m;
      double xs[7] = {0, 0, 0, 0, 1.0, 1.0, rmax(3.0 + k0d, 1.0)};
      ok = tiny_ipm<7, 2 * RDA_MAX_EDGE + 1>(P, xs);
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
