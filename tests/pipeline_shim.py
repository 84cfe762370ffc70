"""TEST-ONLY emulation of the GPU pipeline (rda_kernels.cu: k_begin/k_su/k_cells/k_finalize)
using the g++ build of the same numerical cores (tests/shim.py).  Lets the orchestration and
arithmetic of the kernels be compared with the oracle on a machine without a GPU."""
import numpy as np
import shim
from rda_planner_b200.rda_solver import pack_obstacles, canonical_polygon_rows


class ShimPipeline:
    def __init__(self, T, car, E, N, accelerated=True, prec_su='d', prec_cell='f', **kw):
        self.T, self.N, self.E = T, N, E
        G, h = canonical_polygon_rows(np.asarray(car.G, float), np.asarray(car.h, float))
        self.G, self.h = G, h
        self.R = G.shape[0]
        self.car = car
        self.prec_su, self.prec_cell = prec_su, prec_cell
        self.P = shim.SuParams(T=T, N=N, dynamics=shim.DYN[car.dynamics], accelerated=int(accelerated), dt=kw.get('dt', 0.1),
                               L=car.wheelbase, umax=(shim.C.c_float * 2)(*car.max_speed),
                               ab=(shim.C.c_float * 2)(*(np.asarray(car.max_acce, float) * kw.get('dt', 0.1))),
                               ws=kw.get('ws', 1), wu=kw.get('wu', 1), slack_gain=kw.get('slack_gain', 8),
                               dmin=kw.get('min_sd', 0.1), dmax=kw.get('max_sd', 1.0), ro1=kw.get('ro1', 200),
                               ro2=kw.get('ro2', 1), max_iter=40)
        self.theta = kw.get('z_theta', 0.5) if accelerated else 1.0
        f32 = np.float32
        self.lam = np.zeros((N, E, T), f32); self.mu = np.zeros((N, self.R, T), f32)
        self.z = np.zeros((N, T), f32); self.xi = np.zeros((2, N, T), f32); self.zeta = np.zeros((N, T), f32)
        self.dis = np.ones(T, f32); self.coef = np.zeros((5, N, T), f32); self.pref = np.zeros((2, T), f32)
        self.paths = {}
        self.trace = []

    def solve(self, nom_s, nom_u, ref, ref_speed, obstacles, iters):
        T, N, E = self.T, self.N, self.E
        A, b, kind, count, tv = pack_obstacles(obstacles, T, N, E)
        cur_s = np.asarray(nom_s, np.float32).copy(); cur_u = np.asarray(nom_u, np.float32).copy()
        ref = np.asarray(ref, np.float32)
        for it in range(iters):
            s, u, d, st, nit = shim.su(self.P, cur_s, cur_u, ref, np.float32(ref_speed), self.dis, self.coef[0], self.coef[1],
                                       self.coef[2], self.coef[3], self.coef[4], self.pref, prec=self.prec_su)
            self.su_status = st
            cur_s = s.astype(np.float32); cur_u = u.astype(np.float32); self.dis = d.astype(np.float32)
            hm2 = 0.0; dual = 0.0
            if count > 0:
                for o in range(N):
                    for t in range(T):
                        tc = t + 1 if tv else 0
                        r = shim.cell(self.G, self.h, int(kind[o]), A[o, tc], b[o, tc], cur_s[0:2, t + 1].astype(float),
                                      float(cur_s[2, t]), float(self.dis[t]), float(self.zeta[o, t]),
                                      self.xi[:, o, t].astype(float), self.P.ro2, self.theta, prec=self.prec_cell)
                        self.paths[r['path']] = self.paths.get(r['path'], 0) + 1
                        if r['path'] == 5:
                            dual = np.inf
                            self.failed = (o, t, it)
                            self.failed_args = (int(kind[o]), A[o, tc].copy(), b[o, tc].copy(), cur_s[0:2, t + 1].astype(float), float(cur_s[2, t]), float(self.dis[t]), float(self.zeta[o, t]), self.xi[:, o, t].astype(float))
                            continue
                        lam = r['lam'].astype(np.float32); mu = r['mu'].astype(np.float32)
                        dual += np.sum((lam - self.lam[o, :, t]) ** 2) + np.sum((mu - self.mu[o, :, t]) ** 2) + (np.float32(r['z']) - self.z[o, t]) ** 2
                        self.lam[o, :, t] = lam; self.mu[o, :, t] = mu; self.z[o, t] = r['z']
                        self.zeta[o, t] = r['zeta_new']; self.xi[0, o, t] = r['xi0']; self.xi[1, o, t] = r['xi1']
                        hm2 += r['hm0'] ** 2 + r['hm1'] ** 2
                        self.coef[:, o, t] = [r['ax'], r['ay'], r['c0'], r['gx'], r['gy']]
                    self.pref = cur_s[0:2, 1:].copy()
                resi_pri, resi_dual = np.sqrt(hm2), dual / N
            else:
                resi_pri = resi_dual = 0.0
            self.trace.append((cur_s.copy(), cur_u.copy(), resi_dual, resi_pri))
        return cur_u.astype(float), cur_s.astype(float), resi_dual, resi_pri
