"""Factor-graph container, linearisation and normal equations (oracle; test infrastructure).

Restates what gtsam::NonlinearFactorGraph::linearize / error do for the factor
kinds CGraphGT builds (gtsam/gtsam_graph.cpp:320-368 priors, :630-695 between,
:370-448 projection + point priors, :1118-1298 planes; drivers
gtsam/test_vro_imu_graph.cpp:191-196 CombinedImuFactor).  graph.error = 1/2 sum r^T Omega r (A.1).

Global tangent layout: [poses 6 | vels 3 | biases 6 | planes 3 | points 3].
"""
import copy
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from . import lie
from . import factors as F

CAL_SR4K = (250.5773, 250.5773, 0.0, 90.0, 70.0, -0.8466, 0.5370, 0.0, 0.0)  # gtsam_graph.cpp:544


def _empty(n, *shape):
    return np.zeros((n,) + shape)


class Graph:
    def __init__(self):
        self.R = _empty(0, 3, 3); self.t = _empty(0, 3)
        self.vel = _empty(0, 3); self.bias = _empty(0, 6)
        self.plane = _empty(0, 4); self.point = _empty(0, 3)
        self.f = {}          # factor groups
        self.K = CAL_SR4K
        self.Rs = np.eye(3); self.ts = np.zeros(3)   # body_P_sensor (mp_u2c)
        self.chart = 0                               # Pose3 / Rot3 charts (oracle/lie.py: EXPMAP, FIRST_ORDER_EXPMAP, FIRST_ORDER_CAYLEY)

    # ---- sizes / offsets
    @property
    def dims(self):
        P, Nv, Nb, Npl, L = len(self.R), len(self.vel), len(self.bias), len(self.plane), len(self.point)
        o_v = 6 * P; o_b = o_v + 3 * Nv; o_pl = o_b + 6 * Nb; o_pt = o_pl + 3 * Npl
        return dict(P=P, Nv=Nv, Nb=Nb, Npl=Npl, L=L, o_v=o_v, o_b=o_b, o_pl=o_pl, o_pt=o_pt, n=o_pt + 3 * L)

    def copy(self):
        g = Graph()
        g.__dict__.update({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()})
        g.f = self.f          # factors are immutable, share
        return g

    # ---- retraction (Values::retract, A.1)
    def retract(self, delta):
        d = self.dims
        g = self.copy()
        if d['P']:
            g.R, g.t = lie.pose_retract(self.R, self.t, delta[:d['o_v']].reshape(-1, 6), self.chart)
        g.vel = self.vel + delta[d['o_v']:d['o_b']].reshape(-1, 3)
        g.bias = self.bias + delta[d['o_b']:d['o_pl']].reshape(-1, 6)
        if d['Npl']:
            g.plane = F.plane_retract(self.plane, delta[d['o_pl']:d['o_pt']].reshape(-1, 3))
        g.point = self.point + delta[d['o_pt']:].reshape(-1, 3)
        return g

    # ---- linearisation: list of (r, info, [(offset array, J), ...])
    def linearize(self, jac=True):
        d = self.dims
        out = []
        f = self.f
        if 'prior_pose' in f:
            q = f['prior_pose']; i = q['i']
            res = F.prior_pose(self.R[i], self.t[i], q['R'], q['t'], jac, self.chart)
            out.append((res[0], q['info'], [(6 * i, res[1])]) if jac else (res, q['info'], None))
        if 'prior_vel' in f:
            q = f['prior_vel']; i = q['i']
            res = F.prior_vec(self.vel[i], q['mean'], jac)
            out.append((res[0], q['info'], [(d['o_v'] + 3 * i, res[1])]) if jac else (res, q['info'], None))
        if 'prior_bias' in f:
            q = f['prior_bias']; i = q['i']
            res = F.prior_vec(self.bias[i], q['mean'], jac)
            out.append((res[0], q['info'], [(d['o_b'] + 6 * i, res[1])]) if jac else (res, q['info'], None))
        if 'prior_point' in f:
            q = f['prior_point']; i = q['i']
            res = F.prior_vec(self.point[i], q['mean'], jac)
            out.append((res[0], q['info'], [(d['o_pt'] + 3 * i, res[1])]) if jac else (res, q['info'], None))
        if 'between' in f:
            q = f['between']; i, j = q['i'], q['j']
            res = F.between_pose(self.R[i], self.t[i], self.R[j], self.t[j], q['R'], q['t'], jac, self.chart)
            out.append((res[0], q['info'], [(6 * i, res[1]), (6 * j, res[2])]) if jac else (res, q['info'], None))
        if 'imu' in f:
            q = f['imu']
            pi, vi, pj, vj, bi, bj = q['pi'], q['vi'], q['pj'], q['vj'], q['bi'], q['bj']
            res = F.imu_combined(self.R[pi], self.t[pi], self.vel[vi], self.R[pj], self.t[pj], self.vel[vj],
                                 self.bias[bi], self.bias[bj], q['pim'], jac)
            if jac:
                J = res[1]
                out.append((res[0], q['info'], [(6 * pi, J[0]), (d['o_v'] + 3 * vi, J[1]), (6 * pj, J[2]),
                                                (d['o_v'] + 3 * vj, J[3]), (d['o_b'] + 6 * bi, J[4]),
                                                (d['o_b'] + 6 * bj, J[5])]))
            else:
                out.append((res, q['info'], None))
        if 'proj' in f:
            q = f['proj']; i, l = q['i'], q['l']
            if 'cal' in q:
                # several (Cal3DS2, body_P_sensor) pairs in one graph: q['cal'] indexes self.cals per factor
                n = len(i)
                r = np.zeros((n, 2)); Jp = np.zeros((n, 2, 6)); Jl = np.zeros((n, 2, 3))
                for cidx, (Kc, Rsc, tsc) in enumerate(self.cals):
                    m = q['cal'] == cidx
                    if not m.any():
                        continue
                    rr = F.projection(self.R[i[m]], self.t[i[m]], self.point[l[m]], q['uv'][m], Kc, Rsc, tsc, jac)
                    if jac:
                        r[m], Jp[m], Jl[m] = rr
                    else:
                        r[m] = rr
                res = (r, Jp, Jl) if jac else r
            else:
                res = F.projection(self.R[i], self.t[i], self.point[l], q['uv'], self.K, self.Rs, self.ts, jac)
            info = np.broadcast_to(np.eye(2) / q['sigma'] ** 2, (len(i), 2, 2))
            out.append((res[0], info, [(6 * i, res[1]), (d['o_pt'] + 3 * l, res[2])]) if jac else (res, info, None))
        if 'plane' in f:
            q = f['plane']; i, l = q['i'], q['l']
            res = F.plane_factor(self.R[i], self.t[i], self.plane[l], q['meas'], jac)
            out.append((res[0], q['info'], [(6 * i, res[1]), (d['o_pl'] + 3 * l, res[2])]) if jac else (res, q['info'], None))
        return out

    def error(self):
        """graph.error(values) = 1/2 sum r^T Omega r  (CGraphGT::error, gtsam_graph.cpp:173-176)."""
        e = 0.0
        for r, info, _ in self.linearize(jac=False):
            e += 0.5 * np.einsum('ni,nij,nj->', r, info, r)
        return float(e)

    def normal_equations(self):
        """H = sum J^T Omega J (CSC), g = sum J^T Omega r, err."""
        n = self.dims['n']
        rows, cols, vals = [], [], []
        g = np.zeros(n)
        err = 0.0
        for r, info, blocks in self.linearize(jac=True):
            err += 0.5 * np.einsum('ni,nij,nj->', r, info, r)
            wr = np.einsum('nij,nj->ni', info, r)
            for oa, Ja in blocks:
                da = Ja.shape[-1]
                ga = np.einsum('nia,ni->na', Ja, wr)
                np.add.at(g, (oa[:, None] + np.arange(da)[None, :]).ravel(), ga.ravel())
                WJa = np.einsum('nij,nja->nia', info, Ja)
                for ob, Jb in blocks:
                    db = Jb.shape[-1]
                    Hab = np.einsum('nia,nib->nab', WJa, Jb)
                    rr = (oa[:, None, None] + np.arange(da)[None, :, None]) + np.zeros((1, 1, db), dtype=np.int64)
                    cc = (ob[:, None, None] + np.arange(db)[None, None, :]) + np.zeros((1, da, 1), dtype=np.int64)
                    rows.append(rr.ravel()); cols.append(cc.ravel()); vals.append(Hab.ravel())
        H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsc()
        return H, g, float(err)

    def linearized_error(self, delta):
        """GaussianFactorGraph::error(delta) = 1/2 sum |J delta + r|^2_Omega (undamped)."""
        e = 0.0
        for r, info, blocks in self.linearize(jac=True):
            v = r.copy()
            for o, J in blocks:
                dd = delta[o[:, None] + np.arange(J.shape[-1])[None, :]]
                v = v + np.einsum('nia,na->ni', J, dd)
            e += 0.5 * np.einsum('ni,nij,nj->', v, info, v)
        return float(e)


def solve_direct(H, g, lam):
    """(H + lam I) delta = -g by sparse LU on the full system."""
    n = H.shape[0]
    A = (H + lam * sp.identity(n, format='csc')).tocsc()
    lu = spla.splu(A)
    return lu.solve(-g)


def solve_schur(H, g, lam, n_r):
    """Same system, eliminating the trailing (point) block by Schur complement:
    S = U' - W V'^-1 W^T;  S d_r = -(g_r - W V'^-1 g_l);  d_l = -V'^-1 (g_l + W^T d_r)."""
    n = H.shape[0]
    A = (H + lam * sp.identity(n, format='csc')).tocsc()
    U = A[:n_r, :n_r]; W = A[:n_r, n_r:]; V = A[n_r:, n_r:].tocsc()
    # V is block diagonal 3x3: invert blockwise
    L = (n - n_r) // 3
    Vd = np.zeros((L, 3, 3))
    Vc = V.tocoo()
    Vd[Vc.row // 3, Vc.row % 3, Vc.col % 3] = Vc.data
    Vi = np.linalg.inv(Vd)
    rr = (3 * np.arange(L)[:, None, None] + np.arange(3)[None, :, None]) + np.zeros((1, 1, 3), dtype=np.int64)
    cc = (3 * np.arange(L)[:, None, None] + np.arange(3)[None, None, :]) + np.zeros((1, 3, 1), dtype=np.int64)
    Vinv = sp.coo_matrix((Vi.ravel(), (rr.ravel(), cc.ravel())), shape=(3 * L, 3 * L)).tocsc()
    WVi = (W @ Vinv).tocsc()
    S = (U - WVi @ W.T).tocsc()
    rhs = -(g[:n_r] - WVi @ g[n_r:])
    dr = spla.splu(S).solve(rhs)
    dl = -(Vinv @ (g[n_r:] + W.T @ dr))
    return np.concatenate([dr, dl])
