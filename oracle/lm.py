"""Levenberg-Marquardt control (oracle; test infrastructure).

optimize_gtsam restates gtsam::LevenbergMarquardtOptimizer with DEFAULT parameters, as
invoked by CGraphGT::optimizeGraphBatch (gtsam/gtsam_graph.cpp:1784-1788): SURVEY.md A.7.
  lambda0=1e-5, factor 10 (fixed), lambdaUpperBound 1e5, lambdaLowerBound 0,
  diagonalDamping=false (adds lambda*I), minModelFidelity 1e-3, maxIterations 100,
  relativeErrorTol 1e-5, absoluteErrorTol 1e-5, errorTol 0.
optimize_g2o restates g2o OptimizationAlgorithmLevenberg as driven by
CGraphG2O::optimizeGraph (g2o/g2o_graph.cpp:241-252): SURVEY.md A.8.
PARITY UNPINNED for both (no golden vectors exist in the reference).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from . import lie
from . import factors as F
from .graph import solve_direct, solve_schur


class LMParams:
    def __init__(self, **kw):
        self.lambda_initial = 1e-5
        self.lambda_factor = 10.0
        self.lambda_upper = 1e5
        self.lambda_lower = 0.0
        self.min_model_fidelity = 1e-3
        self.max_iterations = 100
        self.relative_error_tol = 1e-5
        self.absolute_error_tol = 1e-5
        self.error_tol = 0.0
        self.force_iterations = False   # benchmark mode: ignore the convergence test
        self.__dict__.update(kw)


def lm_iterate(graph, lam, params, err, solver='direct', trace=None, it=0):
    """One LevenbergMarquardtOptimizer::iterate(): linearise once, try lambdas. Returns (graph, lam, err)."""
    H, g, err_lin = graph.normal_equations()
    n_r = graph.dims['o_pt']
    while True:
        ok = True
        try:
            if solver == 'schur' and graph.dims['L'] > 0:
                delta = solve_schur(H, g, lam, n_r)
            else:
                delta = solve_direct(H, g, lam)
            if not np.all(np.isfinite(delta)):
                ok = False
        except RuntimeError:
            ok = False
        step_ok = False
        stop = False
        new_err = np.inf
        new_graph = None
        if ok:
            new_lin = graph.linearized_error(delta)
            lin_change = err - new_lin
            if lin_change >= 0:
                new_graph = graph.retract(delta)
                new_err = new_graph.error()
                cost_change = err - new_err
                if lin_change > 1e-20:
                    fidelity = cost_change / lin_change
                    step_ok = fidelity > params.min_model_fidelity
                else:
                    stop = True
                if abs(cost_change) < params.relative_error_tol * err:
                    stop = True
        if trace is not None:
            trace.append(dict(iter=it, lam=lam, err=err, new_err=new_err, accepted=bool(step_ok)))
        if step_ok:
            graph = new_graph
            err = new_err
            lam = max(params.lambda_lower, lam / params.lambda_factor)
            break
        elif not stop:
            lam *= params.lambda_factor
            if lam >= params.lambda_upper:
                break
        else:
            break
    return graph, lam, err


def optimize_gtsam(graph, params=None, solver='direct'):
    """NonlinearOptimizer::defaultOptimize loop around LM iterate. Returns (graph, report)."""
    p = params or LMParams()
    err = graph.error()
    lam = p.lambda_initial
    trace = []
    it = 0
    while True:
        cur = err
        graph, lam, err = lm_iterate(graph, lam, p, err, solver, trace, it)
        it += 1
        if it >= p.max_iterations or not np.isfinite(err):
            break
        if not p.force_iterations and check_convergence(p, cur, err):
            break
    return graph, dict(iterations=it, error=err, lam=lam, trace=trace)


def check_convergence(p, cur, new):
    if p.error_tol >= new:
        return True
    absdec = cur - new
    reldec = absdec / cur if cur != 0 else 0.0
    return (p.relative_error_tol and reldec <= p.relative_error_tol) or (absdec <= p.absolute_error_tol)


# --------------------------------------------------------------------------------------- g2o path
class PoseGraphG2O:
    """SE3 pose graph with g2o semantics (A.8): EdgeSE3 error [t, q_xyz], chi2 = sum e^T Omega e,
    vertex 0 fixed (g2o/g2o_graph.cpp:80-94), oplus X <- X * fromVectorMQT(d)."""

    def __init__(self, R, t, ei, ej, Rm, tm, info, fixed=(0,)):
        self.R, self.t = np.array(R, dtype=np.float64), np.array(t, dtype=np.float64)
        self.ei, self.ej = np.asarray(ei), np.asarray(ej)
        self.Rm, self.tm, self.info = np.asarray(Rm), np.asarray(tm), np.asarray(info)
        self.fixed = np.zeros(len(self.R), dtype=bool); self.fixed[list(fixed)] = True

    def errors(self, R=None, t=None):
        R = self.R if R is None else R; t = self.t if t is None else t
        return F.g2o_edge_se3(R[self.ei], t[self.ei], R[self.ej], t[self.ej], self.Rm, self.tm)

    def chi2(self, R=None, t=None):
        e = self.errors(R, t)
        return float(np.einsum('ni,nij,nj->', e, self.info, e))

    def jacobians(self, eps=None):
        """Jacobians of the edge error w.r.t. the oplus perturbation of each endpoint, at zero perturbation (g2o's analytic
        EdgeSE3 Jacobians are the exact derivative of the same map).  Closed form (E = A B, A = Z^-1, B = X1^-1 X2,
        q_E = (w, v), Q = w I + [v]x):  J2 = [[R_E, 0], [0, Q]],  J1 = [[-R_A, 2 R_A [t_B]x], [0, -Q R_B^T]];
        with eps: central differences through g2o_oplus instead (tests/test_oracle_g2o.py checks one against the other)."""
        if eps is None:
            Ri, ti, Rj, tj = self.R[self.ei], self.t[self.ei], self.R[self.ej], self.t[self.ej]
            Rb, tb = lie.pose_between(Ri, ti, Rj, tj)
            Re, te = lie.pose_between(self.Rm, self.tm, Rb, tb)
            q = lie.quat_from_rot(Re)
            Q = q[:, 0, None, None] * np.eye(3)[None] + lie.skew(q[:, 1:])
            Ra = np.swapaxes(self.Rm, -1, -2)
            n = len(self.ei)
            Ji = np.zeros((n, 6, 6)); Jj = np.zeros((n, 6, 6))
            Jj[:, :3, :3] = Re; Jj[:, 3:, 3:] = Q
            Ji[:, :3, :3] = -Ra; Ji[:, :3, 3:] = 2.0 * Ra @ lie.skew(tb); Ji[:, 3:, 3:] = -Q @ np.swapaxes(Rb, -1, -2)
            return Ji, Jj
        n = len(self.ei)
        Ji = np.zeros((n, 6, 6)); Jj = np.zeros((n, 6, 6))
        Ri, ti, Rj, tj = self.R[self.ei], self.t[self.ei], self.R[self.ej], self.t[self.ej]
        for k in range(6):
            d = np.zeros(6); d[k] = eps
            Rp, tp = F.g2o_oplus(Ri, ti, d); Rq, tq = F.g2o_oplus(Ri, ti, -d)
            Ji[:, :, k] = (F.g2o_edge_se3(Rp, tp, Rj, tj, self.Rm, self.tm) - F.g2o_edge_se3(Rq, tq, Rj, tj, self.Rm, self.tm)) / (2 * eps)
            Rp, tp = F.g2o_oplus(Rj, tj, d); Rq, tq = F.g2o_oplus(Rj, tj, -d)
            Jj[:, :, k] = (F.g2o_edge_se3(Ri, ti, Rp, tp, self.Rm, self.tm) - F.g2o_edge_se3(Ri, ti, Rq, tq, self.Rm, self.tm)) / (2 * eps)
        return Ji, Jj

    def build(self):
        n = len(self.R)
        e = self.errors()
        Ji, Jj = self.jacobians()
        rows, cols, vals = [], [], []
        b = np.zeros(6 * n)
        ar = np.arange(6)
        for (oa, Ja) in ((6 * self.ei, Ji), (6 * self.ej, Jj)):
            WJa = np.einsum('nij,nja->nia', self.info, Ja)
            np.add.at(b, (oa[:, None] + ar[None, :]).ravel(), -np.einsum('nia,ni->na', WJa, e).ravel())
            for (ob, Jb) in ((6 * self.ei, Ji), (6 * self.ej, Jj)):
                Hab = np.einsum('nia,nib->nab', WJa, Jb)
                rows.append(((oa[:, None, None] + ar[None, :, None]) + np.zeros((1, 1, 6), dtype=np.int64)).ravel())
                cols.append(((ob[:, None, None] + ar[None, None, :]) + np.zeros((1, 6, 1), dtype=np.int64)).ravel())
                vals.append(Hab.ravel())
        H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(6 * n, 6 * n)).tocsc()
        free = np.repeat(~self.fixed, 6)
        idx = np.nonzero(free)[0]
        return H[idx][:, idx].tocsc(), b[idx], idx


def optimize_g2o(pg, iterations=20, tau=1e-5, max_trials=10):
    """OptimizationAlgorithmLevenberg::solve repeated `iterations` times (A.8)."""
    lam = None
    ni = 2.0
    trace = []
    n = len(pg.R)
    for it in range(iterations):
        H, b, idx = pg.build()
        cur = pg.chi2()
        if lam is None:
            lam = tau * float(H.diagonal().max())
        rho = 0.0
        trials = 0
        R0, t0 = pg.R.copy(), pg.t.copy()
        while True:
            A = (H + lam * sp.identity(H.shape[0], format='csc')).tocsc()
            dx = spla.splu(A).solve(b)
            full = np.zeros(6 * n); full[idx] = dx
            Rn, tn = F.g2o_oplus(R0, t0, full.reshape(-1, 6))
            new = pg.chi2(Rn, tn)
            scale = float(dx @ (lam * dx + b)) + 1e-3
            rho = (cur - new) / scale
            if rho > 0 and np.isfinite(new):
                alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                lam *= max(1.0 / 3.0, alpha)
                ni = 2.0
                pg.R, pg.t = Rn, tn
                cur = new
            else:
                lam *= ni
                ni *= 2
                if not np.isfinite(lam):
                    break
            trials += 1
            if rho > 0 or trials >= max_trials:
                break
        trace.append(dict(iter=it, chi2=cur, lam=lam, trials=trials))
        if rho <= 0:
            break
    return pg, dict(iterations=len(trace), chi2=pg.chi2(), trace=trace)


def optimize_g2o_calls(pg, iterations=20, per_call=2, tau=1e-5, max_trials=10):
    """CGraphG2O::optimizeGraph as written (g2o/g2o_graph.cpp:241-252): `for (i = 0; i < iter; i += currIt) currIt =
    optimize(ceil(iter/10))` -- every optimize() call restarts OptimizationAlgorithmLevenberg at iteration 0, i.e. lambda
    is re-initialised to tau * max diag(H) and nu to 2 every `per_call` iterations [ext]; an iteration without progress
    (rho == 0, max_trials failures or lambda overflow) terminates that call only."""
    n = len(pg.R)
    trace = []
    done_total, calls = 0, 0
    initial = pg.chi2()
    lam, ni = 0.0, 2.0
    while done_total < iterations:
        done, ok, it = 0, True, 0
        while it < per_call and ok:
            H, b, idx = pg.build()
            cur = pg.chi2()
            if it == 0:
                lam, ni = tau * float(H.diagonal().max()), 2.0
            rho, qmax = 0.0, 0
            R0, t0 = pg.R.copy(), pg.t.copy()
            while True:
                A = (H + lam * sp.identity(H.shape[0], format='csc')).tocsc()
                try:
                    dx = spla.splu(A).solve(b)
                    solved = bool(np.all(np.isfinite(dx)))
                except RuntimeError:
                    dx, solved = np.zeros_like(b), False
                full = np.zeros(6 * n); full[idx] = dx
                Rn, tn = F.g2o_oplus(R0, t0, full.reshape(-1, 6))
                temp = pg.chi2(Rn, tn) if solved else np.finfo(float).max
                scale = (float(dx @ (lam * dx + b)) if solved else 0.0) + 1e-3
                rho = (cur - temp) / scale
                if rho > 0 and np.isfinite(temp):
                    lam *= max(1.0 / 3.0, min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0))
                    ni = 2.0
                    pg.R, pg.t = Rn, tn
                    cur = temp
                else:
                    lam *= ni
                    ni *= 2
                    if not np.isfinite(lam):
                        break
                qmax += 1
                if not (rho < 0 and qmax < max_trials):
                    break
            ok = not (qmax == max_trials or rho == 0 or not np.isfinite(lam))
            done += 1; it += 1
            trace.append(dict(chi2=cur, lam=lam, trials=qmax))
        calls += 1
        done_total += done
        if done == 0:
            break
    return pg, dict(iterations=done_total, calls=calls, initial_chi2=initial, chi2=pg.chi2(), trace=trace)
