"""Host side of the landmark Schur complement (fg_api.cu: build_schur_tables, consumed by k_schur_tiles in fg_schur.cu),
exercised without a GPU through a detached context: a numpy emulation of what the kernel does with the tables -- per tile
and chunk, AND the landmark masks of a row pose and a column pose and count the hits -- must visit every pair of
observations of a common landmark exactly once, and the (start, mask) entries must address the right pose-major records."""
import numpy as np
import pytest
from graph_slam_b200 import abi, synth


def tables(spec, landmark_slice=None):
    ctx = abi.Context(device=-1)
    P = spec['n_poses']
    pims = (abi.Pim * (P - 1))()
    for i in range(P - 1):
        pims[i].dt = 0.1; pims[i].cov[:] = np.eye(15).ravel().tolist()
    abi.load_spec(ctx, spec, preintegrated=pims, landmark_slice=landmark_slice)
    ctx.symbolic(0)
    hdr = ctx.symbolic(41)
    t = dict(ch=int(hdr[0]), n_pairs=int(hdr[2]), tiles=ctx.symbolic(42).reshape(-1, 4), pc_lo=ctx.symbolic(43), pc_n=ctx.symbolic(44),
             pc_ptr=ctx.symbolic(45), pc_ent=ctx.symbolic(46).reshape(-1, 2))
    assert len(t['tiles']) == hdr[1]
    ctx.close()
    return t


@pytest.mark.parametrize('scale,sl,dup', [(0.03, None, 0), (0.05, None, 0), (0.05, (0.3, 0.8), 0), (0.03, None, 60)])
def test_tiles_visit_every_observation_pair_once(fglib, scale, sl, dup):
    spec = synth.make_config('C4', seed=5, scale=scale)
    if dup:
        # extra projection factors on (pose, landmark) pairs that already have one: the tables hold one PRIMARY record per
        # pair (the extras sit behind the primaries of their pose and take no part in the tile products)
        pick = np.random.default_rng(3).choice(len(spec['proj_pose']), size=dup, replace=False)
        pick = np.concatenate([pick, pick[:9]])
        for k in ('proj_pose', 'proj_point', 'proj_uv'):
            spec[k] = np.concatenate([spec[k], spec[k][pick]])
    L = len(spec['point_init']); P = spec['n_poses']
    lo, hi = (0, L) if sl is None else (int(sl[0] * L), int(sl[1] * L))
    t = tables(spec, None if sl is None else (lo, hi))
    CH = t['ch']
    m = (spec['proj_point'] >= lo) & (spec['proj_point'] < hi)
    pose = spec['proj_pose'][m].astype(np.int64); pt = spec['proj_point'][m].astype(np.int64) - lo      # local landmark ids
    # expected: common landmarks of every pose pair, and the pose-major order (by pose, then by landmark)
    vis = np.zeros((P, hi - lo), dtype=bool); vis[pose, pt] = True
    common = vis.astype(np.int64) @ vis.T.astype(np.int64)
    assert t['n_pairs'] == int(sum(k * (k + 1) // 2 for k in vis.sum(0)))
    # pose-major order: per pose its primaries sorted by landmark, then the extra factors
    pm_pose, pm_pt = [], []
    for p in range(P):
        mine = np.sort(pt[pose == p])
        prim = np.unique(mine)
        pm_pose += [p] * len(mine); pm_pt += prim.tolist() + [-1] * (len(mine) - len(prim))
    # entries: for every pose and chunk, mask == landmarks seen in the chunk, start == first pose-major position
    for p in range(P):
        seen = np.nonzero(vis[p])[0]
        if len(seen) == 0:
            assert t['pc_n'][p] == 0
            continue
        assert t['pc_lo'][p] == seen[0] // CH and t['pc_n'][p] == seen[-1] // CH - seen[0] // CH + 1
        for r in range(t['pc_n'][p]):
            c = t['pc_lo'][p] + r
            start, mask = (int(v) for v in t['pc_ent'][t['pc_ptr'][p] + r])
            bits = [b for b in range(CH) if mask >> b & 1]
            assert bits == [int(l - c * CH) for l in seen if l // CH == c]
            for n, b in enumerate(bits):
                assert pm_pose[start + n] == p and pm_pt[start + n] == c * CH + b
    # tiles: accumulate hits the way the kernel does
    got = np.zeros((P, P), dtype=np.int64)
    seen_tiles = set()
    for gi, gj, cb, ce in t['tiles'].tolist():
        assert gj <= gi and (gi, gj) not in seen_tiles
        seen_tiles.add((gi, gj))
        for c in range(cb, ce):
            def mask_of(p):
                if p >= P:
                    return 0
                r = c - t['pc_lo'][p]
                return int(t['pc_ent'][t['pc_ptr'][p] + r][1]) if 0 <= r < t['pc_n'][p] else 0
            rows = [mask_of(gi * 16 + i) for i in range(16)]
            cols = [mask_of(gj * 16 + j) for j in range(16)]
            for i in range(16):
                for j in range(16):
                    if gi == gj and j >= i:
                        continue                      # a diagonal tile holds each unordered pair once; p == q is k_schur_rhs's
                    h = bin(rows[i] & cols[j]).count('1')
                    if h:
                        got[gi * 16 + i, gj * 16 + j] += h
    expect = np.tril(common, -1)
    assert np.array_equal(got, expect)
    # heaviest tiles first
    w = t['tiles'][:, 3] - t['tiles'][:, 2]
    assert np.all(np.diff(w) <= 0)


def test_tile_kernel_index_arithmetic():
    """Python mirror of the index arithmetic of k_schur_tiles (fg_schur.cu): (i) the compute-role lane mapping covers the 256
    blocks of a tile once, a quarter-warp holds 8 distinct row poses and 8 distinct column poses mod 8 (the condition for
    conflict-free 16-byte reads in the [slot][pose][48 B] layout) and the warps of sub-tiles that share a row or a column of the
    tile sit on different schedulers; (ii) in a diagonal tile the parity split gives every landmark of every unordered pose
    pair to exactly one lane; (iii) the staging-role mapping copies every (pose, record, 16-byte part) exactly once, for both
    tile kinds; (iv) slots and hit lists built the kernel's way address the records both poses hold."""
    # (i) lane mapping
    seen = np.zeros((16, 16), dtype=int)
    for warp in range(8):
        sub = (warp >> 1) if (warp >> 1) < 2 else 5 - (warp >> 1)
        for q in range(4):
            rows, cols = set(), set()
            for ii in range(8):
                lane = 8 * q + ii
                dgn = 4 * (warp & 1) + (lane >> 3)
                i = 8 * (sub >> 1) + (lane & 7); j = 8 * (sub & 1) + (((lane & 7) + dgn) & 7)
                seen[i, j] += 1
                rows.add(i % 8); cols.add(j % 8)
                # bank group of a 16-byte read of part 0..2 of any slot: (3 * pose + part) mod 8
            assert len(rows) == 8 and len(cols) == 8
            assert len({(3 * r) % 8 for r in rows}) == 8
    assert np.all(seen == 1)
    sched = {}
    for warp in range(8):
        sub = (warp >> 1) if (warp >> 1) < 2 else 5 - (warp >> 1)
        sched.setdefault(sub, set()).add(warp % 4)
    for a, b in ((0, 1), (2, 3), (0, 2), (1, 3)):                 # sub-tiles sharing a row / a column of the tile
        assert sched[a] | sched[b] == {0, 1, 2, 3}
    # (ii) parity split of a diagonal tile
    for i in range(16):
        for j in range(16):
            if i == j:
                continue
            sel = 0x55555555 if i > j else 0xaaaaaaaa
            other = 0x55555555 if j > i else 0xaaaaaaaa
            assert sel ^ other == 0xffffffff
    # (iii) staging role
    KMAX = 56
    for diag in (False, True):
        npw, st_nr = (2, 5) if diag else (4, 2)
        cnt = np.random.default_rng(1).integers(0, KMAX + 1, size=32)
        copies = {}
        for warp in range(8):
            for lane in range(32):
                st_t, st_part, st_r = lane % npw, (lane // npw) % 3, lane // (3 * npw)
                s = npw * warp + st_t
                my_cnt = cnt[s] if st_r < st_nr else 0
                for kk in range(st_r, my_cnt, st_nr):
                    key = (s, kk, st_part)
                    copies[key] = copies.get(key, 0) + 1
        n_st = 16 if diag else 32
        assert all(v == 1 for v in copies.values())
        assert len(copies) == 3 * int(cnt[:n_st].sum())
    # (iv) slots and hit lists over three words
    rng = np.random.default_rng(2)
    for _ in range(200):
        own_i = [int(x) for x in rng.integers(0, 2 ** 32, size=3)]; own_j = [int(x) for x in rng.integers(0, 2 ** 32, size=3)]
        other_of_i = [int(x) for x in rng.integers(0, 2 ** 32, size=3)]; other_of_j = [int(x) for x in rng.integers(0, 2 ** 32, size=3)]
        fi = [own_i[w] & (other_of_i[w] | own_j[w]) for w in range(3)]      # f = own & OR(other side); the other side contains j
        fj = [own_j[w] & (other_of_j[w] | own_i[w]) for w in range(3)]
        slots_i = [(w, b) for w in range(3) for b in range(32) if (fi[w] >> b) & 1]   # staged records of pose i in slot order
        slots_j = [(w, b) for w in range(3) for b in range(32) if (fj[w] >> b) & 1]
        pi = pj = 0
        hits = []
        for w in range(3):
            hit = fi[w] & fj[w]
            while hit:
                b = (hit & -hit).bit_length() - 1
                hit &= hit - 1
                low = (1 << b) - 1
                hits.append((pi + bin(fi[w] & low).count('1'), pj + bin(fj[w] & low).count('1'), w, b))
            pi += bin(fi[w]).count('1'); pj += bin(fj[w]).count('1')
        want = [(w, b) for w in range(3) for b in range(32) if ((own_i[w] & own_j[w]) >> b) & 1]
        assert [(w, b) for _, _, w, b in hits] == want                     # every common landmark, once, in landmark order
        for a, bq, w, b in hits:
            assert slots_i[a] == (w, b) and slots_j[bq] == (w, b)        # both slots hold that landmark's records
