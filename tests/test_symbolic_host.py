"""Host logic of the reduced solve, exercised without a GPU through a detached context (fg_create(-1,..)):
the symbolic structure (ordering, supernodes, row lists, update lists) produced by fg_symbolic.cpp is used by a
numpy emulation of the left-looking supernodal algorithm that k_chol / k_backsolve implement, and the result is
compared with a dense solve of the oracle's reduced system."""
import numpy as np
import pytest
from graph_slam_b200 import abi, synth
from oracle import build


def reduced_system(spec, lam):
    """Oracle: damped reduced system (Schur complement over points) and the oracle->reduced index map."""
    g = build.from_spec(spec)
    H, grad, err = g.normal_equations()
    d = g.dims
    n_r = d['o_pt']
    H = H.toarray() + lam * np.eye(H.shape[0])
    U, W, V = H[:n_r, :n_r], H[:n_r, n_r:], H[n_r:, n_r:]
    if d['L']:
        Vi = np.linalg.inv(V)
        S = U - W @ Vi @ W.T
        b = -(grad[:n_r] - W @ Vi @ grad[n_r:])
    else:
        S, b = U, -grad[:n_r]
    return g, S, b


def perm_from_offsets(ctx, g):
    d = g.dims
    off = {k: ctx.symbolic(w) for k, w in (('pose', 11), ('vel', 12), ('bias', 13), ('plane', 14))}
    perm = np.zeros(d['o_pt'], dtype=np.int64)       # oracle reduced index -> solver reduced index
    for i in range(d['P']):
        perm[6 * i:6 * i + 6] = off['pose'][i] + np.arange(6)
    for i in range(d['Nv']):
        perm[d['o_v'] + 3 * i: d['o_v'] + 3 * i + 3] = off['vel'][i] + np.arange(3)
    for i in range(d['Nb']):
        perm[d['o_b'] + 6 * i: d['o_b'] + 6 * i + 6] = off['bias'][i] + np.arange(6)
    for i in range(d['Npl']):
        perm[d['o_pl'] + 3 * i: d['o_pl'] + 3 * i + 3] = off['plane'][i] + np.arange(3)
    return perm


def emulate(sym, Sp, bp):
    """numpy emulation of k_chol (left-looking, rhs row) + k_backsolve on the panel layout."""
    n_r, n_sn = int(sym[0][0]), int(sym[0][1])
    col0, ncols, nrows, rowptr, valptr, rowidx, uptr, ud, ua, ub = sym[1:11]
    aug = np.zeros((n_r + 1, n_r + 1)); aug[:n_r, :n_r] = Sp; aug[n_r, :n_r] = bp
    panels = []
    for s in range(n_sn):
        rows = rowidx[rowptr[s]:rowptr[s] + nrows[s]]
        cols = np.arange(col0[s], col0[s] + ncols[s])
        assert np.array_equal(rows[:ncols[s]], cols) and rows[-1] == n_r and np.all(np.diff(rows) > 0)
        # structure check: every nonzero of the lower part of these columns lies in the row list
        full = np.nonzero(np.abs(aug[col0[s]:, cols]).sum(1))[0] + col0[s]
        assert set(full.tolist()) <= set(rows.tolist()), 'S has an entry outside the symbolic structure'
        panels.append(np.tril(aug[np.ix_(rows, cols)][:ncols[s]], 0).tolist() and aug[np.ix_(rows, cols)].copy())
        panels[s][:ncols[s]] = np.tril(panels[s][:ncols[s]])
    for s in range(n_sn):
        rows_s = rowidx[rowptr[s]:rowptr[s] + nrows[s]]
        pos = {int(r): k for k, r in enumerate(rows_s)}
        for u in range(uptr[s], uptr[s + 1]):
            d, a, b = int(ud[u]), int(ua[u]), int(ub[u])
            assert d < s
            rows_d = rowidx[rowptr[d]:rowptr[d] + nrows[d]]
            Ld = panels[d]
            upd = Ld[a:] @ Ld[a:b].T
            for ii, R in enumerate(rows_d[a:]):
                assert int(R) in pos, 'descendant row missing from ancestor structure'
                for jj in range(b - a):
                    C = int(rows_d[a + jj])
                    if R >= C:
                        panels[s][pos[int(R)], C - col0[s]] -= upd[ii, jj]
        nc = ncols[s]
        Ldd = np.linalg.cholesky(panels[s][:nc] + np.tril(panels[s][:nc], -1).T)
        panels[s][:nc] = Ldd
        panels[s][nc:] = np.linalg.solve(Ldd, panels[s][nc:].T).T
    x = np.zeros(n_r)
    for s in range(n_sn - 1, -1, -1):
        rows = rowidx[rowptr[s]:rowptr[s] + nrows[s]]
        nc = ncols[s]
        t = panels[s][-1] - panels[s][nc:-1].T @ x[rows[nc:-1]]
        x[col0[s]:col0[s] + nc] = np.linalg.solve(panels[s][:nc].T, t)
    return x


def emulate_fronts(ctx, sym, Sp, bp):
    """numpy emulation of the leaf-front path (fg_front.cu + the phase split of k_chol_rs): leaf members use the
    reduced update lists, every leaf's contribution to the outside is one dense matrix U = sum_d A_d A_d^T gathered
    through the position maps, and the remaining supernodes subtract the U entries that fall in their panels."""
    n_r, n_sn = int(sym[0][0]), int(sym[0][1])
    col0, ncols, nrows, rowptr, valptr, rowidx = sym[1:7]
    sn_leaf = ctx.symbolic(21)
    uptr, ud, ua, ub = (ctx.symbolic(w) for w in (22, 23, 24, 25))
    fr_rowptr, fr_rows, members = ctx.symbolic(26), ctx.symbolic(27), ctx.symbolic(28).reshape(-1, 2)
    tf_ptr, tf_leaf, pm_ptr, posmap = ctx.symbolic(29), ctx.symbolic(30), ctx.symbolic(31), ctx.symbolic(32)
    sched_a, sched_c = ctx.symbolic(34), ctx.symbolic(35)
    assert sorted(sched_a.tolist() + sched_c.tolist()) == list(range(n_sn))
    assert all(sn_leaf[s] >= 0 for s in sched_a) and all(sn_leaf[s] < 0 for s in sched_c)
    aug = np.zeros((n_r + 1, n_r + 1)); aug[:n_r, :n_r] = Sp; aug[n_r, :n_r] = bp
    rows_of = [rowidx[rowptr[s]:rowptr[s] + nrows[s]] for s in range(n_sn)]
    panels = []
    for s in range(n_sn):
        cols = np.arange(col0[s], col0[s] + ncols[s])
        P = aug[np.ix_(rows_of[s], cols)].copy(); P[:ncols[s]] = np.tril(P[:ncols[s]])
        panels.append(P)
    done = np.zeros(n_sn, dtype=bool)

    def finish(s):
        pos = {int(r): k for k, r in enumerate(rows_of[s])}
        for u in range(uptr[s], uptr[s + 1]):
            d, a, b = int(ud[u]), int(ua[u]), int(ub[u])
            assert done[d], 'reduced update list is not topological within its phase'
            Ld = panels[d]
            upd = Ld[a:] @ Ld[a:b].T
            for ii, R in enumerate(rows_of[d][a:]):
                for jj in range(b - a):
                    C = int(rows_of[d][a + jj])
                    if R >= C:
                        panels[s][pos[int(R)], C - col0[s]] -= upd[ii, jj]
        nc = ncols[s]
        Ldd = np.linalg.cholesky(panels[s][:nc] + np.tril(panels[s][:nc], -1).T)
        panels[s][:nc] = Ldd
        panels[s][nc:] = np.linalg.solve(Ldd, panels[s][nc:].T).T
        done[s] = True

    for s in sched_a:
        finish(int(s))
    U = []
    for l, (lo, hi) in enumerate(members):
        Rl = fr_rows[fr_rowptr[l]:fr_rowptr[l + 1]]
        Ul = np.zeros((len(Rl), len(Rl)))
        for d in range(lo, hi):
            pm = posmap[pm_ptr[d]:pm_ptr[d + 1]]
            assert len(pm) == len(Rl)
            for k, pidx in enumerate(pm):
                assert pidx < 0 or rows_of[d][pidx] == Rl[k]
            A = np.where(pm[:, None] >= 0, panels[d][np.maximum(pm, 0)], 0.0)
            # every row of d outside its leaf must be in the front
            outside = set(int(r) for r in rows_of[d][ncols[d]:] if r >= col0[hi - 1] + ncols[hi - 1])
            assert outside <= set(int(r) for r in Rl)
            Ul += A @ A.T
        U.append((Rl, Ul))
    for s in sched_c:
        s = int(s)
        for e in range(tf_ptr[s], tf_ptr[s + 1]):
            Rl, Ul = U[int(tf_leaf[e])]
            idx = {int(r): k for k, r in enumerate(Rl)}
            for rr, g in enumerate(rows_of[s]):
                if int(g) not in idx:
                    continue
                for c in range(ncols[s]):
                    gc = int(col0[s] + c)
                    if gc in idx and g >= gc:
                        panels[s][rr, c] -= Ul[idx[int(g)], idx[gc]]
        finish(s)
    x = np.zeros(n_r)
    for s in range(n_sn - 1, -1, -1):
        rows = rows_of[s]; nc = ncols[s]
        t = panels[s][-1] - panels[s][nc:-1].T @ x[rows[nc:-1]]
        x[col0[s]:col0[s] + nc] = np.linalg.solve(panels[s][:nc].T, t)
    return x


def emulate_rowsplit(ctx, sym, Sp, bp):
    """numpy emulation of k_chol_rs (fg_chol_rs.cu): the unit of work is a block of <= 128 panel rows of a supernode.
    A unit reads its ASSEMBLED rows and applies every descendant update restricted to the descendant rows the host mapped
    onto its rows (rs_map / rs_colinv).  The head unit (rows [0, min(nr, 128)): the diagonal block and the rows right below
    it) factors the diagonal block, then solves its other rows; the further row blocks solve theirs against that factor."""
    n_r, n_sn = int(sym[0][0]), int(sym[0][1])
    col0, ncols, nrows, rowptr, valptr, rowidx = sym[1:7]
    use_fr = bool(ctx.symbolic(33)[0])
    ok, n_a, n_units = (int(v) for v in ctx.symbolic(39)[:3])
    assert ok
    units = ctx.symbolic(36).reshape(-1, 4); moff = ctx.symbolic(37); rmap = ctx.symbolic(38)
    colinv = ctx.symbolic(40).reshape(-1, 32)
    assert len(units) == n_units
    uptr, ud, ua, ub = (ctx.symbolic(w) for w in ((22, 23, 24, 25) if use_fr else (7, 8, 9, 10)))
    # the kernel's own list: every update of the list in use, cut into column slices of <= 16
    qptr = ctx.symbolic(47); qrec = ctx.symbolic(48).reshape(-1, 5)       # (d, source update, k0, K, 8-column group mask)
    assert len(colinv) == len(qrec) and np.all(ncols <= 32)
    # ... in the order the descendants' units are handed out (a unit pulls its updates in list order)
    first_unit = {}
    for k, un in enumerate(units.tolist()):
        first_unit.setdefault(un[0], k)
    for s in range(n_sn):
        us = sorted(range(uptr[s], uptr[s + 1]), key=lambda u: first_unit[int(ud[u])])
        want = [(int(ud[u]), u, k0, min(16, int(ncols[ud[u]]) - k0)) for u in us for k0 in range(0, int(ncols[ud[u]]), 16)]
        assert [tuple(r[:4]) for r in qrec[qptr[s]:qptr[s + 1]].tolist()] == want
    aug = np.zeros((n_r + 1, n_r + 1)); aug[:n_r, :n_r] = Sp; aug[n_r, :n_r] = bp
    rows_of = [rowidx[rowptr[s]:rowptr[s] + nrows[s]] for s in range(n_sn)]
    A0, Lf = [], []                      # assembled panels (read-only) and factored panels
    for s in range(n_sn):
        cols = np.arange(col0[s], col0[s] + ncols[s])
        P = aug[np.ix_(rows_of[s], cols)].copy(); P[:ncols[s]] = np.tril(P[:ncols[s]])
        A0.append(P); Lf.append(np.full_like(P, np.nan))
    arrived = np.zeros(n_sn, dtype=int)
    covered = [np.zeros(nrows[s], dtype=int) for s in range(n_sn)]
    U = None

    def fronts():
        fr_rowptr, fr_rows, members = ctx.symbolic(26), ctx.symbolic(27), ctx.symbolic(28).reshape(-1, 2)
        pm_ptr, posmap = ctx.symbolic(31), ctx.symbolic(32)
        out = []
        for l, (lo, hi) in enumerate(members):
            Rl = fr_rows[fr_rowptr[l]:fr_rowptr[l + 1]]
            Ul = np.zeros((len(Rl), len(Rl)))
            for d in range(lo, hi):
                assert arrived[d] == -1
                pm = posmap[pm_ptr[d]:pm_ptr[d + 1]]
                A = np.where(pm[:, None] >= 0, Lf[d][np.maximum(pm, 0)], 0.0)
                Ul += A @ A.T
            out.append((Rl, Ul))
        return out

    for k, (s, r0, r1, nblk) in enumerate(units.tolist()):
        if use_fr and k == n_a:
            U = fronts()
        nc = int(ncols[s])
        diag = r0 == 0
        assert (r1 == min(int(nrows[s]), 128) if diag else 128 <= r0 < r1 <= nrows[s]) and r1 - r0 <= 128
        loc = list(range(r0, r1))
        P = A0[s][loc].copy()
        g = [int(rows_of[s][i]) for i in loc]
        if use_fr and k >= n_a:
            tf_ptr, tf_leaf = ctx.symbolic(29), ctx.symbolic(30)
            for e in range(tf_ptr[s], tf_ptr[s + 1]):
                Rl, Ul = U[int(tf_leaf[e])]
                idx = {int(r): q for q, r in enumerate(Rl)}
                for rr, gr in enumerate(g):
                    if gr not in idx:
                        continue
                    for c in range(nc):
                        gc = int(col0[s] + c)
                        if gc in idx and gr >= gc:
                            P[rr, c] -= Ul[idx[gr], idx[gc]]
        for q, qq in enumerate(range(qptr[s], qptr[s + 1])):
            d, u, k0, K, mask = (int(v) for v in qrec[qq])
            a, b = int(ua[u]), int(ub[u])
            assert arrived[d] == -1, 'descendant not complete when its update is pulled'
            Ld = Lf[d][:, k0:k0 + K]
            nloc = len(loc)
            mp = rmap[moff[k] + q * nloc: moff[k] + (q + 1) * nloc]
            where = {int(r): i for i, r in enumerate(rows_of[d]) if i >= a}
            assert mp.tolist() == [where.get(gr, a - 1) - a for gr in g], 'row map does not name the descendant rows'
            ci = colinv[qq]
            assert ci.tolist() == [where.get(int(col0[s] + c), a - 1) - a if c < nc else -1 for c in range(32)]
            assert all(0 <= v < b - a for v in ci if v >= 0) and sum(v >= 0 for v in ci) == b - a
            assert mask == sum(1 << n for n in range(4) if any(ci[8 * n + e] >= 0 for e in range(8)))
            Bfull = np.zeros((Ld.shape[1], 32))
            for c in range(32):
                if ci[c] >= 0 and (mask >> (c // 8)) & 1:
                    Bfull[:, c] = Ld[a + ci[c]]
            for lr in range(nloc):
                if mp[lr] >= 0:
                    upd = Ld[a + mp[lr]] @ Bfull
                    for c in range(nc):
                        if c <= r0 + lr:            # lower triangle of the diagonal block, every column of the rows below it
                            P[lr, c] -= upd[c]
        if diag:
            assert arrived[s] == 0, 'the head unit comes first'
            D = np.tril(P[:nc])
            Lf[s][:nc] = np.linalg.cholesky(D + np.tril(D, -1).T)
            Lf[s][nc:r1] = np.linalg.solve(Lf[s][:nc], P[nc:].T).T
        else:
            assert arrived[s] >= 1, 'the diagonal factor is not there yet'
            Lf[s][r0:r1] = np.linalg.solve(Lf[s][:nc], P.T).T
        covered[s][r0:r1] += 1
        arrived[s] += 1
        if arrived[s] == nblk:
            arrived[s] = -1                  # every unit's done flag is up
    assert all(np.all(c == 1) for c in covered) and np.all(arrived == -1)
    x = np.zeros(n_r)
    for s in range(n_sn - 1, -1, -1):
        rows = rows_of[s]; nc = ncols[s]
        t = Lf[s][-1] - Lf[s][nc:-1].T @ x[rows[nc:-1]]
        x[col0[s]:col0[s] + nc] = np.linalg.solve(Lf[s][:nc].T, t)
    return x


@pytest.mark.parametrize('name,scale', [('C1', 1.0), ('C2', 0.08), ('C3', 0.08), ('C4', 0.03), ('C4', 0.06)])
def test_symbolic_and_left_looking(fglib, name, scale):
    spec = synth.make_config(name, seed=2, scale=scale)
    lam = 1e-3
    g, S, b = reduced_system(spec, lam)
    ctx = abi.Context(device=-1)
    # detached context cannot preintegrate on a device: hand it oracle-preintegrated records
    pims = None
    if 'imu' in g.f:
        q = g.f['imu']['pim']
        pims = (abi.Pim * len(q['dt']))()
        for i in range(len(q['dt'])):
            pims[i].dt = q['dt'][i]
            pims[i].preint[:] = q['preint'][i].tolist(); pims[i].H_ba[:] = q['Hba'][i].ravel().tolist()
            pims[i].H_bg[:] = q['Hbg'][i].ravel().tolist(); pims[i].bias_hat[:] = q['bias_hat'][i].tolist()
            pims[i].cov[:] = q['cov'][i].ravel().tolist(); pims[i].gravity[:] = q['gravity'].tolist()
    abi.load_spec(ctx, spec, preintegrated=pims)
    sym = [ctx.symbolic(w) for w in range(11)]
    perm = perm_from_offsets(ctx, g)
    n_r = len(perm)
    assert sym[0][0] == n_r and sorted(perm.tolist()) == list(range(n_r))
    Sp = np.zeros_like(S); Sp[np.ix_(perm, perm)] = S
    bp = np.zeros_like(b); bp[perm] = b
    x = emulate(sym, Sp, bp)
    # schedule = topological order by level; ancestor lists = transpose of the update lists
    sched, level = ctx.symbolic(15), ctx.symbolic(16)
    anc_ptr, anc_t, anc_a, anc_b = (ctx.symbolic(w) for w in (17, 18, 19, 20))
    n_sn = int(sym[0][1])
    assert sorted(sched.tolist()) == list(range(n_sn)) and np.all(np.diff(level[sched]) >= 0)
    uptr, ud, ua, ub = sym[7:11]
    pairs_u = set()
    for t in range(n_sn):
        for u in range(uptr[t], uptr[t + 1]):
            assert level[ud[u]] < level[t]
            pairs_u.add((int(ud[u]), t, int(ua[u]), int(ub[u])))
    pairs_a = set()
    for d in range(n_sn):
        ents = [(int(anc_t[e]), int(anc_a[e]), int(anc_b[e])) for e in range(anc_ptr[d], anc_ptr[d + 1])]
        assert [e[1] for e in ents] == sorted(e[1] for e in ents)
        if ents:
            assert ents[0][1] == sym[2][d] and ents[-1][2] == sym[3][d] - 1
        for (t, a, b) in ents:
            pairs_a.add((d, t, a, b))
    assert pairs_u == pairs_a
    ref = np.linalg.solve(Sp, bp)
    assert np.allclose(x, ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())
    if ctx.symbolic(33)[0]:
        xf = emulate_fronts(ctx, sym, Sp, bp)
        assert np.allclose(xf, ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())
    if ctx.symbolic(39)[0]:
        xr = emulate_rowsplit(ctx, sym, Sp, bp)
        assert np.allclose(xr, ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())
    ctx.close()


def test_detached_context_has_no_cpu_solver(fglib):
    ctx = abi.Context(device=-1)
    ctx.add_pose(abi.symbol('x', 0), np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]))
    with pytest.raises(abi.FgError) as e:
        ctx.error()
    assert e.value.code == -4
    with pytest.raises(abi.FgError):
        ctx.optimize()
    with pytest.raises(abi.FgError) as e:
        ctx.marginal_covariance(abi.symbol('x', 0))
    assert e.value.code == -4
    with pytest.raises(abi.FgError) as e:
        ctx.marginal_covariance(abi.symbol('x', 7))            # unknown key: ValuesKeyDoesNotExist
    assert e.value.code == -3
    ctx.close()
