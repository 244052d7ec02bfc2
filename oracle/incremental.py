"""ISAM2 semantics as CGraphGT uses them, restated (oracle; test infrastructure, never imported by the product).

CGraphGT::optimizeGraphIncremental (gtsam/gtsam_graph.cpp:1768-1776):  isam2->update(*mp_new_fac, *mp_new_node);
*mp_node_values = isam2->calculateEstimate();  with ISAM2Params relinearizeThreshold = 0.1, relinearizeSkip = 1
(initISAM2Params, gtsam_graph.cpp:93-99) -- what test_vro_imu_graph.cpp:344 and test_ba_imu_graph.cpp:427 call once per
frame.  GTSAM's ISAM2 [ext] keeps a linearisation point theta and a delta (estimate = theta (+) delta).  One update():
  1. new variables enter with theta = their initial value, delta = 0;
  2. "fluid relinearisation": every variable whose delta has a component >= relinearizeThreshold in magnitude moves its
     linearisation point, theta_j <- theta_j (+) delta_j (checked every relinearizeSkip updates);
  3. the factors touching new / relinearised variables are re-linearised at theta and the affected part of the Bayes tree
     is re-eliminated; delta is recomputed by back-substitution.
With wildfire threshold 0 step 3 is the solution of the full linear system  J(theta)^T Omega J(theta) delta = -J^T Omega r(theta)
(a factor that touches no moved variable linearises to the same thing it did before), i.e. ONE undamped Gauss-Newton
solve at theta -- which is what is restated here and what fg_update_incremental runs on the device."""
import numpy as np
from . import lie
from . import factors as F
from .graph import Graph, solve_direct, solve_schur


def local(theta, est):
    """delta with est = theta.retract(delta), per variable block (the inverse of Graph.retract)."""
    d = theta.dims
    out = np.zeros(d['n'])
    if d['P']:
        out[:d['o_v']] = lie.pose_local(theta.R, theta.t, est.R, est.t).ravel()
    out[d['o_v']:d['o_b']] = (est.vel - theta.vel).ravel()
    out[d['o_b']:d['o_pl']] = (est.bias - theta.bias).ravel()
    if d['Npl']:
        out[d['o_pl']:d['o_pt']] = F.plane_local(theta.plane, est.plane).ravel()
    out[d['o_pt']:] = (est.point - theta.point).ravel()
    return out


def _blocks(d):
    return (('R', 0, 6, d['P']), ('vel', d['o_v'], 3, d['Nv']), ('bias', d['o_b'], 6, d['Nb']), ('plane', d['o_pl'], 3, d['Npl']),
            ('point', d['o_pt'], 3, d['L']))


class IncrementalSmoother:
    def __init__(self, relinearize_threshold=0.1, relinearize_skip=1):
        self.thr, self.skip = relinearize_threshold, relinearize_skip
        self.theta = None          # Graph at the linearisation point
        self.est = None            # Graph at the estimate
        self.n_updates = 0

    def update(self, graph, solver='direct'):
        """`graph`: ALL factors so far; its values are used only for variables this smoother has not seen (new ones are
        appended at the end of each variable type).  Returns dict(error_before, error_after, n_relinearized)."""
        new = graph.copy()
        if self.theta is not None:
            for name in ('R', 't', 'vel', 'bias', 'plane', 'point'):
                old_t, old_e = getattr(self.theta, name), getattr(self.est, name)
                th = getattr(new, name).copy(); es = th.copy()
                th[:len(old_t)] = old_t; es[:len(old_e)] = old_e
                setattr(new, name, th)
                setattr(self, '_e_' + name, es)
            est = new.copy()
            for name in ('R', 't', 'vel', 'bias', 'plane', 'point'):
                setattr(est, name, getattr(self, '_e_' + name))
        else:
            est = new.copy()
        theta = new
        self.n_updates += 1
        n_rel = 0
        if self.n_updates % self.skip == 0:
            dl = local(theta, est)
            d = theta.dims
            for name, off, dim, cnt in _blocks(d):
                if cnt == 0:
                    continue
                move = np.abs(dl[off:off + dim * cnt].reshape(cnt, dim)).max(1) >= self.thr
                n_rel += int(move.sum())
                if name == 'R':
                    theta.R[move] = est.R[move]; theta.t[move] = est.t[move]
                else:
                    getattr(theta, name)[move] = getattr(est, name)[move]
        H, g, e0 = theta.normal_equations()
        n_r = theta.dims['o_pt']
        delta = solve_schur(H, g, 0.0, n_r) if (solver == 'schur' and theta.dims['L']) else solve_direct(H, g, 0.0)
        self.theta = theta
        self.est = theta.retract(delta)
        return dict(error_before=e0, error_after=self.est.error(), n_relinearized=n_rel)
