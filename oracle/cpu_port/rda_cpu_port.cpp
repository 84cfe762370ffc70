// TEST / BASELINE INFRASTRUCTURE (oracle/): CPU build of the numerical cores of
// rda_planner_b200/csrc (the same templates the CUDA kernels instantiate), compiled with g++.
//   * shim_cell_*, shim_su_*: single sub-problem entry points, used by tests/ to compare the kernel
//     arithmetic with the numpy oracle on a machine without a GPU;
//   * port_solve_batch: the whole ADMM loop of rda_kernels.cu (k_begin / k_su / k_cells /
//     k_finalize orchestration, float32 state, cold start) over a batch of instances with OpenMP
//     over instances — the compiled multi-core CPU baseline bench.py reports ("port").
// Never loaded by the product (rda_planner_b200/, RDA_planner/).
#include <vector>
#include <cstring>
#include <cmath>
#include <cstdlib>
#include <omp.h>
#include "../../rda_planner_b200/csrc/cell_solver.cuh"
#include "../../rda_planner_b200/csrc/cell_disc_robot.cuh"
#include "../../rda_planner_b200/csrc/su_solver.cuh"
#include "../../rda_planner_b200/csrc/cell_lean.cuh"
#include "../../rda_planner_b200/csrc/cell_lean2.cuh"
#ifdef RDA_CELL_STATS
static void rda_coh_stat(int it, int lean_ok, int coh_ok);
#endif

using namespace rda;

#ifdef RDA_SOC_STATS
static long long g_soc[2];
extern "C" void rda_soc_stat(int newton) {
#pragma omp atomic
  g_soc[0] += 1;
#pragma omp atomic
  g_soc[1] += newton;
}
static long long g_dr[4];
extern "C" void rda_dr_stat(int what) {
#pragma omp atomic
  g_dr[what & 3] += 1;
}
extern "C" void port_dr_stats(long long* out) { for (int i = 0; i < 4; ++i) { out[i] = g_dr[i]; g_dr[i] = 0; } }
extern "C" void port_soc_stats(long long* out, int reset) { out[0] = g_soc[0]; out[1] = g_soc[1]; if (reset) g_soc[0] = g_soc[1] = 0; }
#endif

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

// disc body (cone_type 'norm2'): h = (cx, cy, -r)
template <typename Real>
static int cell_dr_impl(const float* h, int kind, int E, const float* A, const float* b, double px, double py, double phi,
                        double dbar, double zeta, double xi0, double xi1, double ro2, double theta, double* out, int forms = 1) {
  const float Gd[6] = {1.f, 0.f, 0.f, 1.f, 0.f, 0.f};
  RobotGeom rb;
  int rc = robot_geom_from_halfspaces(Gd, h, 3, &rb, RDA_ROBOT_DISC);
  if (rc) return rc;
  CellOut<Real> o;
  cell_solve_dr<Real>(rb, kind, E, A, b, (Real)px, (Real)py, (Real)cos(phi), (Real)sin(phi), (Real)dbar,
                      (Real)zeta, (Real)xi0, (Real)xi1, (Real)ro2, (Real)theta, o, forms != 0);
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
  size_t bytes = su_work_bytes<Real>(T, N);
  std::vector<char> buf(bytes + 64);
  char* base = (char*)(((uintptr_t)buf.data() + 63) & ~(uintptr_t)63);
  su_work_layout<Real>(T, N, &W, base);
  for (int i = 0; i < 3 * (T + 1); ++i) { W.lins[i] = (Real)lins[i]; W.ref[i] = (Real)ref[i]; }
  for (int i = 0; i < 2 * T; ++i) { W.linu[i] = (Real)linu[i]; W.pref[i] = (Real)pref[i]; }
  for (int i = 0; i < T; ++i) W.d[i] = (Real)dis[i];
  for (int i = 0; i < N * T; ++i) { W.hx[i] = hx[i]; W.hy[i] = hy[i]; W.hc[i] = hc[i]; }
  W.vref = (Real)vref;
  SeqCtx ctx;
  int st = su_solve<Real, Real, SeqCtx>(*P, W, ctx, gx, gy, iters);
  for (int i = 0; i < 3 * (T + 1); ++i) s[i] = W.s[i];
  for (int i = 0; i < 2 * T; ++i) u[i] = W.u[i];
  for (int i = 0; i < T; ++i) d[i] = W.d[i];
  return st;
}

template <int EC, int RC>
static int lean_impl(const float* G, const float* h, int R, int kind, int E, const float* A, const float* b, double px,
                     double py, double phi, double dbar, double zeta, double xi0, double xi1, double theta, double* out) {
  RobotGeom rb;
  int rc = robot_geom_from_halfspaces(G, h, R, &rb);
  if (rc) return rc;
  LeanOut<EC, RC> o;
  bool ok = cell_lean<EC, RC>(rb, kind, E, A, b, (float)px, (float)py, (float)cos(phi), (float)sin(phi), (float)dbar,
                              (float)zeta, (float)xi0, (float)xi1, (float)theta, o);
  for (int i = 0; i < 28; ++i) out[i] = 0;
  out[27] = ok ? 0 : 6;
  if (!ok) return 0;
  for (int i = 0; i < EC; ++i) out[i] = o.lam[i];
  for (int j = 0; j < RC; ++j) out[8 + j] = o.mu[j];
  out[16] = o.z; out[17] = o.zeta_new; out[18] = 0; out[19] = 0; out[20] = o.ax; out[21] = o.ay; out[22] = o.c0;
  out[23] = o.gx; out[24] = o.gy; out[25] = o.feat; out[26] = 0;
  return 0;
}

// coherent first pass (cell_lean2.cuh): out as lean_impl, out[25] = new feature byte, out[27] = 6 when declined
extern "C" int shim_cell_lean2_4(const float* G, const float* h, int R, int E, const float* A, const float* b, int feat,
                                 double px, double py, double phi, double dbar, double zeta, double theta, double* out) {
  RobotGeom rb;
  int rc = robot_geom_from_halfspaces(G, h, R, &rb);
  if (rc) return rc;
  RobotAux ra;
  robot_aux_from_geom(rb, &ra);
  ObstacleGeom<4> og;
  obstacle_geometry<4>(E, A, b, og);
  LeanOut<4, 4> o;
  int nf = cell_lean2<4, 4>(rb, ra, og, feat, (float)px, (float)py, (float)cos(phi), (float)sin(phi), (float)dbar,
                            (float)zeta, (float)theta, o);
  for (int i = 0; i < 28; ++i) out[i] = 0;
  out[27] = nf >= 0 ? 0 : 6;
  if (nf < 0) return 0;
  for (int i = 0; i < 4; ++i) out[i] = o.lam[i];
  for (int j = 0; j < 4; ++j) out[8 + j] = o.mu[j];
  out[16] = o.z; out[17] = o.zeta_new; out[20] = o.ax; out[21] = o.ay; out[22] = o.c0;
  out[23] = o.gx; out[24] = o.gy; out[25] = nf;
  return 0;
}

extern "C" {
int shim_cell_lean4(const float* G, const float* h, int R, int kind, int E, const float* A, const float* b, double px,
                    double py, double phi, double dbar, double zeta, double xi0, double xi1, double ro2, double theta,
                    double* out) {
  (void)ro2;
  return lean_impl<4, 4>(G, h, R, kind, E, A, b, px, py, phi, dbar, zeta, xi0, xi1, theta, out);
}
int shim_cell_lean8(const float* G, const float* h, int R, int kind, int E, const float* A, const float* b, double px,
                    double py, double phi, double dbar, double zeta, double xi0, double xi1, double ro2, double theta,
                    double* out) {
  (void)ro2;
  return lean_impl<8, 8>(G, h, R, kind, E, A, b, px, py, phi, dbar, zeta, xi0, xi1, theta, out);
}
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
int shim_cell_dr_d(const float* h, int kind, int E, const float* A, const float* b, double px, double py, double phi,
                   double dbar, double zeta, double xi0, double xi1, double ro2, double theta, double* out) {
  return cell_dr_impl<double>(h, kind, E, A, b, px, py, phi, dbar, zeta, xi0, xi1, ro2, theta, out);
}
// searched closed forms off: every cell that is not a plain inactive one goes through the two-cone barrier programmes
int shim_cell_dr_barrier_d(const float* h, int kind, int E, const float* A, const float* b, double px, double py, double phi,
                           double dbar, double zeta, double xi0, double xi1, double ro2, double theta, double* out) {
  return cell_dr_impl<double>(h, kind, E, A, b, px, py, phi, dbar, zeta, xi0, xi1, ro2, theta, out, 0);
}
int shim_cell_dr_f(const float* h, int kind, int E, const float* A, const float* b, double px, double py, double phi,
                   double dbar, double zeta, double xi0, double xi1, double ro2, double theta, double* out) {
  return cell_dr_impl<float>(h, kind, E, A, b, px, py, phi, dbar, zeta, xi0, xi1, ro2, theta, out);
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


static long long g_su_hist[64];
extern "C" void port_su_hist(long long* out, int reset) { for (int i = 0; i < 64; ++i) { out[i] = g_su_hist[i]; if (reset) g_su_hist[i] = 0; } }

extern "C" int shim_su_batched_fwd(const SuParams* Pin, int nb, float* cur_s, float* cur_u, const float* ref_s, const float* pref,
                                   const float* coef, float* dis, const float* ref_speed, const int* done, int* status, int* iters,
                                   int* counters, int max_iter);

// ------------------------------------------------------------------------------------------------
// Whole hot path on the CPU (mirrors rda_kernels.cu; see the kernel comments for reference lines).
// Layouts as in include/rda_b200.h.  Every instance starts cold (constructor state).
// ------------------------------------------------------------------------------------------------
extern "C" int port_solve_batch(const rda_config* cfg, const rda_tunables* tun, int B, const float* nom_s,
                                const float* nom_u, const float* ref_s, const float* ref_speed,
                                const float* obs_A, const float* obs_b, const int* obs_kind,
                                const int* obs_count, int tv, int iter_num, float thr, float* u_opt,
                                float* s_opt, float* resi_pri, float* resi_dual, int* iters_out, int* fails_out, int nthreads) {
  RobotGeom rb;
  int rc = robot_geom_from_halfspaces(cfg->G, cfg->h, cfg->robot_edges, &rb, cfg->robot_cone);
  if (rc) return rc;
  const int T = cfg->receding, N = cfg->max_obs_num, E = cfg->max_edge_num, R = cfg->robot_edges, NT = N * T;
  SuParams P;
  P.T = T; P.N = N; P.dynamics = cfg->dynamics; P.accelerated = cfg->accelerated;
  P.dt = cfg->step_time; P.L = cfg->wheelbase;
  P.umax[0] = cfg->max_speed[0]; P.umax[1] = cfg->max_speed[1];
  P.ab[0] = cfg->acce_bound[0]; P.ab[1] = cfg->acce_bound[1];
  P.ws = cfg->ws; P.wu = cfg->wu; P.slack_gain = tun->slack_gain; P.dmin = tun->min_sd; P.dmax = tun->max_sd;
  P.ro1 = tun->ro1; P.ro2 = tun->ro2; P.max_iter = 40;
  P.mu0 = getenv("RDA_PORT_MU0") ? (float)atof(getenv("RDA_PORT_MU0")) : 1.0f;
  P.prune = getenv("RDA_PORT_PRUNE") ? (float)atof(getenv("RDA_PORT_PRUNE")) : 0.5f;      // as the library default
  const float theta = cfg->accelerated ? tun->z_theta : 1.0f;
  if (nthreads > 0) omp_set_num_threads(nthreads);
  const size_t bytes = su_work_bytes<double>(T, N);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    std::vector<char> buf(bytes + 64);
    char* base = (char*)(((uintptr_t)buf.data() + 63) & ~(uintptr_t)63);
    SuWork<double> W;
    su_work_layout<double>(T, N, &W, base);
    std::vector<float> lam((size_t)N * E * T, 0.f), mu((size_t)N * R * T, 0.f), z(NT, 0.f), xi(2 * NT, 0.f),
        zeta(NT, 0.f), dis(T, 1.f), coef(5 * (size_t)NT, 0.f), pref(2 * T, 0.f);
#ifdef RDA_CELL_STATS
    std::vector<int> featv(NT > 0 ? NT : 1, 0);
#endif
    const bool use_batched = getenv("RDA_PORT_SU_BATCHED") && atoi(getenv("RDA_PORT_SU_BATCHED")) != 0;
    const bool use_lean2 = getenv("RDA_PORT_LEAN2") && atoi(getenv("RDA_PORT_LEAN2")) != 0;
    std::vector<int> feat2(NT > 0 ? NT : 1, 0);
    std::vector<ObstacleGeom<4>> og2(N > 0 ? N : 1);
    RobotAux ra2;
    robot_aux_from_geom(rb, &ra2);
    long long coh_hits = 0;
    if (use_lean2 && E == 4 && !tv && !rb.disc)
      for (int o = 0; o < N; ++o) obstacle_geometry<4>(E, obs_A + ((size_t)b * N + o) * E * 2, obs_b + ((size_t)b * N + o) * E, og2[o]);
    std::vector<float> cs(nom_s + (size_t)b * 3 * (T + 1), nom_s + (size_t)(b + 1) * 3 * (T + 1));
    std::vector<float> cu(nom_u + (size_t)b * 2 * T, nom_u + (size_t)(b + 1) * 2 * T);
    const float* rf = ref_s + (size_t)b * 3 * (T + 1);
    float rp = 0.f, rd = 0.f;
    int it = 0, nfail = 0, first_fail = -1, su_iters = 0, su_bad = 0;
    for (it = 0; it < iter_num; ++it) {
      for (int i = 0; i < 3 * (T + 1); ++i) { int r = i / (T + 1), t = i % (T + 1); W.lins[3 * t + r] = cs[i]; W.ref[3 * t + r] = rf[i]; }
      for (int i = 0; i < 2 * T; ++i) { int r = i / T, t = i % T; W.linu[2 * t + r] = cu[i]; W.pref[2 * t + r] = pref[i]; }
      for (int t = 0; t < T; ++t) W.d[t] = dis[t];
      for (int i = 0; i < NT; ++i) { W.hx[i] = coef[i]; W.hy[i] = coef[NT + i]; W.hc[i] = coef[2 * NT + i]; }
      W.vref = ref_speed[b];
      SeqCtx ctx;
      int nit = 0;
      int st;
      if (use_batched) {
        // the batched pipeline works on the product layouts directly (and writes cs / cu / dis itself)
        int stat = 0, itc = 0, cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dn = 0;
        shim_su_batched_fwd(&P, 1, cs.data(), cu.data(), rf, pref.data(), coef.data(), dis.data(), ref_speed + b, &dn, &stat, &itc, cnt, 28);
        nit = cnt[3];
        st = (stat & RDA_ST_SU_NONFINITE) ? 2 : ((stat & RDA_ST_SU_NOT_CONVERGED) ? 1 : 0);
        for (int i = 0; i < 3 * (T + 1); ++i) { int r = i / (T + 1), t = i % (T + 1); W.s[3 * t + r] = cs[i]; }
        for (int i = 0; i < 2 * T; ++i) { int r = i / T, t = i % T; W.u[2 * t + r] = cu[i]; }
        for (int t = 0; t < T; ++t) W.d[t] = dis[t];
      } else
      st = su_solve<double, double, SeqCtx>(P, W, ctx, coef.data() + 3 * NT, coef.data() + 4 * NT, &nit);
      su_iters += nit;
      if (W.restarts) {
#pragma omp atomic
        g_su_hist[0] += 1;      // slot 0 (never an iteration count): pruned solves that had to be repeated
      }
#pragma omp atomic
      g_su_hist[nit < 63 ? nit : 63] += 1;
      if (st != 0) ++su_bad;
      if (st != 2) {
        for (int i = 0; i < 3 * (T + 1); ++i) { int r = i / (T + 1), t = i % (T + 1); cs[i] = (float)W.s[3 * t + r]; }
        for (int i = 0; i < 2 * T; ++i) { int r = i / T, t = i % T; cu[i] = (float)W.u[2 * t + r]; }
        for (int t = 0; t < T; ++t) dis[t] = (float)W.d[t];
      }
      double hm2 = 0, dual = 0;
      if (N > 0 && obs_count[b] != 0) {
        for (int o = 0; o < N; ++o)
          for (int t = 0; t < T; ++t) {
            int tc = tv ? t + 1 : 0, Tc = tv ? T + 1 : 1;
            size_t ob = ((size_t)b * N + o) * Tc + tc;
            float ph = cs[2 * (T + 1) + t];
#ifdef RDA_CELL_STATS
            // emulation of the coherent first pass (cell_lean2.cuh): how often does the support-vertex pair of
            // the previous ADMM iteration still certify the closest pair?  statistics only, results unused
            if (E == 4 && R == 4 && !tv && obs_kind[(size_t)b * N + o] == RDA_OBS_POLYGON && xi[o * T + t] == 0.f && xi[NT + o * T + t] == 0.f) {
              LeanOut<4, 4> l1, l2;
              const bool ok1 = cell_lean<4, 4>(rb, RDA_OBS_POLYGON, E, obs_A + ob * E * 2, obs_b + ob * E, cs[t + 1], cs[(T + 1) + t + 1],
                                               cosf(ph), sinf(ph), dis[t], zeta[o * T + t], 0.f, 0.f, theta, l1);
              ObstacleGeom<4> og;
              obstacle_geometry<4>(E, obs_A + ob * E * 2, obs_b + ob * E, og);
              RobotAux ra;
              robot_aux_from_geom(rb, &ra);
              const int nf = cell_lean2<4, 4>(rb, ra, og, featv[o * T + t], cs[t + 1], cs[(T + 1) + t + 1], cosf(ph), sinf(ph), dis[t],
                                              zeta[o * T + t], theta, l2);
              rda_coh_stat(it, ok1 ? 1 : 0, nf >= 0 ? 1 : 0);
              featv[o * T + t] = nf >= 0 ? nf : (ok1 ? l1.feat : 0);
            } else {
              rda_coh_stat(it, 0, 0);
              featv[o * T + t] = 0;
            }
#endif
            CellOut<float> out;
            // emulation of the coherent pipeline (RDA_PORT_LEAN2=1): k_cells_coh -> listed cell_lean -> generic solver
            bool lean_done = false;
            if (use_lean2 && !rb.disc && E == 4 && R == 4 && !tv && obs_kind[(size_t)b * N + o] == RDA_OBS_POLYGON) {
              LeanOut<4, 4> lo;
              int nf = -1;
              if (xi[o * T + t] == 0.f && xi[NT + o * T + t] == 0.f)
                nf = cell_lean2<4, 4>(rb, ra2, og2[o], feat2[o * T + t], cs[t + 1], cs[(T + 1) + t + 1], cosf(ph), sinf(ph), dis[t],
                                      zeta[o * T + t], theta, lo);
              if (nf < 0) {
                const bool ok1 = cell_lean<4, 4>(rb, RDA_OBS_POLYGON, E, obs_A + ob * E * 2, obs_b + ob * E, cs[t + 1], cs[(T + 1) + t + 1],
                                                 cosf(ph), sinf(ph), dis[t], zeta[o * T + t], xi[o * T + t], xi[NT + o * T + t], theta, lo);
                nf = ok1 ? lo.feat : -1;
              } else ++coh_hits;
              feat2[o * T + t] = nf >= 0 ? nf : 0;
              if (nf >= 0) {
                lean_done = true;
                for (int i = 0; i < 8; ++i) { out.lam[i] = i < 4 ? lo.lam[i] : 0.f; out.mu[i] = i < 4 ? lo.mu[i] : 0.f; }
                out.z = lo.z; out.zeta_new = lo.zeta_new; out.xi0_new = 0.f; out.xi1_new = 0.f; out.ax = lo.ax; out.ay = lo.ay;
                out.c0 = lo.c0; out.gx = lo.gx; out.gy = lo.gy; out.hm0 = 0.f; out.hm1 = 0.f; out.path = CELL_FAST_INACTIVE;
              }
            }
            if (rb.disc)
              cell_solve_dr<float>(rb, obs_kind[(size_t)b * N + o], E, obs_A + ob * E * 2, obs_b + ob * E, cs[t + 1],
                                   cs[(T + 1) + t + 1], cosf(ph), sinf(ph), dis[t], zeta[o * T + t], xi[o * T + t],
                                   xi[NT + o * T + t], (float)P.ro2, theta, out);
            else if (!lean_done)
            cell_solve<float>(rb, obs_kind[(size_t)b * N + o], E, obs_A + ob * E * 2, obs_b + ob * E, cs[t + 1],
                              cs[(T + 1) + t + 1], cosf(ph), sinf(ph), dis[t], zeta[o * T + t], xi[o * T + t],
                              xi[NT + o * T + t], (float)P.ro2, theta, out);
            if (out.path == CELL_FAILED) { dual = INFINITY; ++nfail; if (first_fail < 0) first_fail = (it * N + o) * T + t; continue; }
            float acc = 0.f;
            for (int i = 0; i < E; ++i) { float nv = out.lam[i], df = nv - lam[((size_t)o * E + i) * T + t]; acc += df * df; lam[((size_t)o * E + i) * T + t] = nv; }
            for (int j = 0; j < R; ++j) { float nv = out.mu[j], df = nv - mu[((size_t)o * R + j) * T + t]; acc += df * df; mu[((size_t)o * R + j) * T + t] = nv; }
            float dz = out.z - z[o * T + t];
            acc += dz * dz;
            z[o * T + t] = out.z;
            dual += acc;
            zeta[o * T + t] = out.zeta_new;
            xi[o * T + t] = out.xi0_new; xi[NT + o * T + t] = out.xi1_new;
            hm2 += out.hm0 * out.hm0 + out.hm1 * out.hm1;
            coef[o * T + t] = out.ax; coef[NT + o * T + t] = out.ay; coef[2 * NT + o * T + t] = out.c0;
            coef[3 * NT + o * T + t] = out.gx; coef[4 * NT + o * T + t] = out.gy;
            if (o == 0) { pref[t] = cs[t + 1]; pref[T + t] = cs[(T + 1) + t + 1]; }
          }
        rp = (float)sqrt(hm2); rd = (float)(dual / N);
      } else { rp = 0.f; rd = 0.f; }
      if (rd < thr && rp < thr) { ++it; break; }
    }
    for (int i = 0; i < 3 * (T + 1); ++i) s_opt[(size_t)b * 3 * (T + 1) + i] = cs[i];
    for (int i = 0; i < 2 * T; ++i) u_opt[(size_t)b * 2 * T + i] = cu[i];
    resi_pri[b] = rp; resi_dual[b] = rd;
    if (iters_out) iters_out[b] = it;
    if (fails_out) { fails_out[4 * b] = nfail; fails_out[4 * b + 1] = first_fail; fails_out[4 * b + 2] = su_iters; fails_out[4 * b + 3] = su_bad; if (use_lean2) fails_out[4 * b + 1] = (int)coh_hits; }
  }
  return 0;
}


// ---- serial launch emulation of the batched su-QP pipeline (rda_planner_b200/csrc/su_batched.cuh) --------
// The kernels of that file have no shared memory and no intra-block synchronisation, so running their
// bodies once per (block, thread) index reproduces the device arithmetic exactly (up to libm vs CUDA math).
namespace emu {
struct Dim { int x = 0, y = 0, z = 0; };
static thread_local Dim blockIdx, threadIdx, blockDim;
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline int atomicSub(int* p, int v) { int o = *p; *p -= v; return o; }
}
#define RDA_SB_EMULATE 1
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
using emu::blockIdx; using emu::threadIdx; using emu::blockDim; using emu::atomicAdd; using emu::atomicSub;
#include "../../rda_planner_b200/csrc/su_batched.cuh"
#undef __global__
#undef __device__
#undef __forceinline__
#undef __launch_bounds__

template <typename F> static void emu_launch(int gx, int gy, int bx, F&& body) {
  emu::blockDim.x = bx;
  for (int y = 0; y < gy; ++y)
    for (int x = 0; x < gx; ++x)
      for (int t = 0; t < bx; ++t) { emu::blockIdx.x = x; emu::blockIdx.y = y; emu::threadIdx.x = t; body(); }
}

// nb instances, product layouts (include/rda_b200.h): cur_s/ref_s [nb][3][T+1], cur_u/pref [nb][2][T],
// coef [nb][5][N][T], dis [nb][T].  Outputs in place (cur_s, cur_u, dis) + status / iters / counters[8].
extern "C" int shim_su_batched(const SuParams* Pin, int nb, float* cur_s, float* cur_u, const float* ref_s,
                               const float* pref, const float* coef, float* dis, const float* ref_speed,
                               const int* done, int* status, int* iters, int* counters, int max_iter) {
  SuParams P = *Pin;
  P.max_iter = max_iter;
  const int T = P.T, N = P.N;
  const size_t bytes = su_batch_layout(nb, T, N, nullptr, nullptr);
  std::vector<char> buf(bytes + 512, 0);
  char* base = (char*)(((uintptr_t)buf.data() + 255) & ~(uintptr_t)255);
  SuBatch W;
  su_batch_layout(nb, T, N, &W, base);
  *W.n_active = 0;
  const double Mrows = 4.0 * T + (N > 0 ? 2.0 * T : 0.0) + 4.0 * (T - 1) + (P.accelerated ? (double)N * T : 0.0);
  SbOut out{cur_s, cur_u, dis, status, iters, counters};
  const int gs = (nb + 127) / 128, gi = (nb + 31) / 32;
  emu_launch(gs, T, 128, [&] { ksb_setup(W, P, cur_s, cur_u, ref_s, pref, coef, dis, ref_speed, done); });
  emu_launch(1, 1, 1, [&] { ksb_compact(W); });
  emu_launch(gi, 1, 32, [&] { ksb_rollout(W); });
  for (int it = 0; it <= P.max_iter; ++it) {
    emu_launch(gs, T, 128, [&] { ksb_assemble(W, P, it); });
    emu_launch(gi, 1, 32, [&] { ksb_riccati<true>(W, P, it, Mrows, out); });
    emu_launch(1, 1, 1, [&] { ksb_compact(W); });
    if (it == P.max_iter) break;
    emu_launch(gs, T, 128, [&] { ksb_steplen<0>(W, P, it); });
    emu_launch(gs, 1, 128, [&] { ksb_reduce<0>(W, Mrows); });
    emu_launch(gs, T, 128, [&] { ksb_corrector(W, P, it); });
    emu_launch(gi, 1, 32, [&] { ksb_riccati<false>(W, P, it, Mrows, out); });
    emu_launch(gs, T, 128, [&] { ksb_steplen<1>(W, P, it); });
    emu_launch(gs, 1, 128, [&] { ksb_reduce<1>(W, Mrows); });
  }
  return *W.n_active;
}

// ---- front end cores (rda_planner_b200/csrc/frontend.cuh), one instance per call -----------------
#include "../../rda_planner_b200/csrc/frontend.cuh"

extern "C" int shim_pre_process(int dynamics, int T, double dt, double L, const float* state, const float* vel,
                                double ref_speed, const float* path, int P, int start_index, double threshold,
                                int ind_range, float* nom_s, float* ref_s) {
  return rda::pre_process_one(dynamics, T, dt, L, state, vel, ref_speed, path, P, start_index, threshold, ind_range,
                              nom_s, ref_s);
}

// Emulates k_convert_obstacles for one instance: M raw shapes -> N slots.
extern "C" int shim_convert_obstacles(int M, int N, int T, int E, double dt, int time_varying, int order, const float* state,
                                      const int* kind, const int* nv, const float* xy, const float* radius,
                                      const float* vel, int count, float* obs_A, float* obs_b, int* obs_kind) {
  if (count > M) count = M;
  std::vector<double> keys(count > 0 ? count : 1);
  for (int j = 0; j < count; ++j)
    keys[j] = order ? rda::obstacle_key(kind[j], nv[j], xy + (size_t)j * RDA_MAX_EDGE * 2, state[0], state[1]) : (double)j;
  const int Tc = time_varying ? T + 1 : 1;
  for (int n = 0; n < N; ++n) {
    float* A = obs_A + (size_t)n * Tc * E * 2;
    float* b = obs_b + (size_t)n * Tc * E;
    const int src = rda::obstacle_slot_source(n, count, 1, keys.data());
    if (src < 0) {
      for (int i = 0; i < Tc * E; ++i) { A[2 * i] = 0.f; A[2 * i + 1] = 0.f; b[i] = 0.f; }
      obs_kind[n] = RDA_OBS_POLYGON;
      continue;
    }
    obs_kind[n] = kind[src];
    for (int t = 0; t < Tc; ++t)
      rda::obstacle_rows(kind[src], nv[src], xy + (size_t)src * RDA_MAX_EDGE * 2, radius[src], vel[2 * src], vel[2 * src + 1], t,
                         dt, E, A + (size_t)t * E * 2, b + (size_t)t * E);
  }
  return count;
}

extern "C" void shim_motion_predict(int dynamics, double dt, double L, const double* s, double v0, double v1, double* out) {
  rda::motion_predict(dynamics, dt, L, s, v0, v1, out);
}

// ---- statistics of the slow cell path (host analysis only; build with -DRDA_CELL_STATS) -------------
#ifdef RDA_CELL_STATS
namespace rda { long long g_cell_stats[8]; }     // which geometric situations reach the slow path (cell_solver.cuh)
static long long g_stat_hist[3][512];
extern "C" void rda_cell_stat(int what, int value) {
  if (what < 0 || what > 2) return;
  if (value < 0) value = 0;
  if (value > 511) value = 511;
#pragma omp atomic
  g_stat_hist[what][value] += 1;
}
static long long g_case[1024];
extern "C" void rda_case_stat(int line) {
  if (line < 0 || line > 1023) return;
#pragma omp atomic
  g_case[line] += 1;
}
extern "C" void port_case_stats(long long* out) { memcpy(out, g_case, sizeof(g_case)); }
static long long g_coh[64][3];
static void rda_coh_stat(int it, int lean_ok, int coh_ok) {
  if (it < 0 || it > 63) return;
#pragma omp atomic
  g_coh[it][0] += 1;
#pragma omp atomic
  g_coh[it][1] += lean_ok;
#pragma omp atomic
  g_coh[it][2] += coh_ok;
}
extern "C" void port_coh_stats(long long* out) { memcpy(out, g_coh, sizeof(g_coh)); }
extern "C" void port_cell_stats(long long* out) { memcpy(out, g_stat_hist, sizeof(g_stat_hist)); }
extern "C" void port_cell_situations(long long* out) { memcpy(out, rda::g_cell_stats, sizeof(rda::g_cell_stats)); }
#endif

extern "C" int shim_su_batched_fwd(const SuParams* Pin, int nb, float* cur_s, float* cur_u, const float* ref_s, const float* pref,
                                   const float* coef, float* dis, const float* ref_speed, const int* done, int* status, int* iters,
                                   int* counters, int max_iter) {
  return shim_su_batched(Pin, nb, cur_s, cur_u, ref_s, pref, coef, dis, ref_speed, done, status, iters, counters, max_iter);
}
