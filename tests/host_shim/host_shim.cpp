// TEST-ONLY: compiles the numerical cores of rda_planner_b200/csrc (the same templates the
// CUDA kernels instantiate) with g++ so that their arithmetic can be checked against the
// oracle on a machine without a GPU.  Not part of the product; never loaded by
// rda_planner_b200.
#include <vector>
#include <cstring>
#include "../../rda_planner_b200/csrc/cell_solver.cuh"
#include "../../rda_planner_b200/csrc/su_solver.cuh"

using namespace rda;

template <typename Real>
static int cell_impl(const float* G, const float* h, int R, int kind, int E, const float* A, const float* b,
                     double px, double py, double phi, double dbar, double zeta, double xi0, double xi1,
                     double ro2, double theta, double* out /* lam[8] mu[8] z zeta_new xi0 xi1 ax ay c0 gx gy hm0 hm1 path */) {
  RobotGeom rb;
  int rc = robot_geom_from_halfspaces(G, h, R, &rb);
  if (rc) return rc;
  CellOut<Real> o;
  cell_solve<Real>(rb, kind, E, A, b, (Real)px, (Real)py, (Real)cos(phi), (Real)sin(phi), (Real)dbar,
                   (Real)zeta, (Real)xi0, (Real)xi1, (Real)ro2, (Real)theta, o);
  for (int i = 0; i < 8; ++i) out[i] = o.lam[i];
  for (int i = 0; i < 8; ++i) out[8 + i] = o.mu[i];
  double tail[] = {(double)o.z, (double)o.zeta_new, (double)o.xi0_new, (double)o.xi1_new, (double)o.ax,
                   (double)o.ay, (double)o.c0, (double)o.gx, (double)o.gy, (double)o.hm0, (double)o.hm1,
                   (double)o.path};
  memcpy(out + 16, tail, sizeof(tail));
  return 0;
}

template <typename Real>
static int su_impl(const SuParams* P, const double* lins, const double* linu, const double* ref, double vref,
                   const double* dis, const float* hx, const float* hy, const float* hc, const float* gx,
                   const float* gy, const double* pref, double* s, double* u, double* d, int* iters) {
  const int T = P->T, N = P->N;
  SuWork<Real> W;
  size_t bytes = su_work_layout<Real>(T, N, nullptr, nullptr);
  std::vector<char> buf(bytes + 64);
  char* base = (char*)(((uintptr_t)buf.data() + 63) & ~(uintptr_t)63);
  su_work_layout<Real>(T, N, &W, base);
  for (int i = 0; i < 3 * (T + 1); ++i) { W.lins[i] = (Real)lins[i]; W.ref[i] = (Real)ref[i]; }
  for (int i = 0; i < 2 * T; ++i) { W.linu[i] = (Real)linu[i]; W.pref[i] = (Real)pref[i]; }
  for (int i = 0; i < T; ++i) W.d[i] = (Real)dis[i];
  for (int i = 0; i < N * T; ++i) { W.hx[i] = hx[i]; W.hy[i] = hy[i]; W.hc[i] = hc[i]; }
  W.vref = (Real)vref;
  SeqCtx ctx;
  int st = su_solve<Real, SeqCtx>(*P, W, ctx, gx, gy, iters);
  for (int i = 0; i < 3 * (T + 1); ++i) s[i] = W.s[i];
  for (int i = 0; i < 2 * T; ++i) u[i] = W.u[i];
  for (int i = 0; i < T; ++i) d[i] = W.d[i];
  return st;
}

extern "C" {
int shim_cell_d(const float* G, const float* h, int R, int kind, int E, const float* A, const float* b,
                double px, double py, double phi, double dbar, double zeta, double xi0, double xi1, double ro2,
                double theta, double* out) {
  return cell_impl<double>(G, h, R, kind, E, A, b, px, py, phi, dbar, zeta, xi0, xi1, ro2, theta, out);
}
int shim_cell_f(const float* G, const float* h, int R, int kind, int E, const float* A, const float* b,
                double px, double py, double phi, double dbar, double zeta, double xi0, double xi1, double ro2,
                double theta, double* out) {
  return cell_impl<float>(G, h, R, kind, E, A, b, px, py, phi, dbar, zeta, xi0, xi1, ro2, theta, out);
}
int shim_su_d(const SuParams* P, const double* lins, const double* linu, const double* ref, double vref,
              const double* dis, const float* hx, const float* hy, const float* hc, const float* gx,
              const float* gy, const double* pref, double* s, double* u, double* d, int* iters) {
  return su_impl<double>(P, lins, linu, ref, vref, dis, hx, hy, hc, gx, gy, pref, s, u, d, iters);
}
int shim_su_f(const SuParams* P, const double* lins, const double* linu, const double* ref, double vref,
              const double* dis, const float* hx, const float* hy, const float* hc, const float* gx,
              const float* gy, const double* pref, double* s, double* u, double* d, int* iters) {
  return su_impl<float>(P, lins, linu, ref, vref, dis, hx, hy, hc, gx, gy, pref, s, u, d, iters);
}
}
