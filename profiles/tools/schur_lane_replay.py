"""Offline replay of k_schur_tiles' schedule on the C5 graph structure (numpy, no GPU): for a sample of tiles over their whole
landmark range, the hits every lane (= one 6x6 block of a 16 x 16 pose tile) has per staging round, under different lane -> block
mappings and round lengths.  A warp runs until its busiest lane is out of hits, so

    lane utilisation (sum over warps) = hits / (32 * sum over rounds and warps of max-lane hits)
    lane utilisation (barrier bound)  = hits / (32 * 8 * sum over rounds of the max over the tile's 256 lanes)

The first is what the DFMA / shared-memory pipes see when other CTAs fill a waiting warp's issue slots, the second what a CTA alone
would get (its warps wait for each other at the round's barrier).  DESIGN.md section 2.1 quotes these numbers; the ncu figure to
compare with is smsp__thread_inst_executed_per_inst_executed on the DFMA lines (10.7 / 32 in round 1, 16-18 / 32 now).

    python profiles/tools/schur_lane_replay.py > profiles/r2_schur_lane_replay.txt        (about two minutes)
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graph_slam_b200 import synth

CH = 32
spec = synth.make_config('C5')
pp, pl = spec['proj_pose'].astype(np.int64), spec['proj_point'].astype(np.int64)
P, L = spec['n_poses'], len(spec['point_init'])
nch = (L + CH - 1) // CH
vis = np.zeros((P, nch), dtype=np.uint32)
np.bitwise_or.at(vis, (pp, pl // CH), (np.uint32(1) << (pl % CH).astype(np.uint32)))
PC = np.array([bin(x).count('1') for x in range(65536)], dtype=np.int32)


def popc(a):
    return PC[a & 0xffff] + PC[a >> 16]


G = P // 16
g_lo = np.array([np.nonzero(vis[g * 16:(g + 1) * 16].any(0))[0].min() for g in range(G)])
g_hi = np.array([np.nonzero(vis[g * 16:(g + 1) * 16].any(0))[0].max() + 1 for g in range(G)])
I, J = np.meshgrid(np.arange(16), np.arange(16), indexing='ij')


def map_wrapped16():            # round 1 / early round 2: half-warp = one wrapped diagonal of the whole tile
    return [[(i, (i + 2 * w + h) & 15) for h in range(2) for i in range(16)] for w in range(8)]


def map_sub8():                 # final: quarter-warp = one wrapped diagonal of an 8 x 8 sub-tile, two warps per sub-tile
    m = []
    for w in range(8):
        s = (w >> 1) if (w >> 1) < 2 else 5 - (w >> 1)
        m.append([(8 * (s >> 1) + ii, 8 * (s & 1) + ((ii + 4 * (w & 1) + g) & 7)) for g in range(4) for ii in range(8)])
    return m


MAPS = {'wrapped diagonals of the 16x16 tile': map_wrapped16(), '8x8 sub-tiles': map_sub8()}
print('C5: %d poses, %d landmarks, %d projections; tiles of row groups 100..111, all their landmark chunks' % (P, L, len(pp)))
print('%-40s %6s %10s %18s %18s' % ('lane mapping', 'words', 'rounds', 'util (sum warps)', 'util (barrier)'))
for NW in (1, 2, 3):
    res = {k: [0, 0, 0, 0] for k in MAPS}
    for gi in range(100, 112):
        for gj in range(max(0, gi - 4), gi + 1):
            cb, ce = max(g_lo[gi], g_lo[gj]), min(g_hi[gi], g_hi[gj])
            if cb >= ce:
                continue
            A = vis[gi * 16:gi * 16 + 16, cb:ce]; B = vis[gj * 16:gj * 16 + 16, cb:ce]
            H = popc(A[:, None, :] & B[None, :, :]).astype(np.int32)
            if gi == gj:                                     # the two lanes of a pose pair split its landmarks by bit parity
                He = popc(A[:, None, :] & B[None, :, :] & np.uint32(0x55555555)); Ho = H - He
                H = np.where((I > J)[:, :, None], He, Ho); H[I == J] = 0
            N = ce - cb
            for r0 in range(0, N, NW):
                Hr = H[:, :, r0:r0 + NW].sum(2)
                if Hr.sum() == 0:
                    continue
                for k, m in MAPS.items():
                    steps = [max(Hr[i, j] for (i, j) in lanes) for lanes in m]
                    res[k][0] += sum(steps); res[k][1] += 8 * max(steps); res[k][2] += int(Hr.sum()); res[k][3] += 1
    for k, (s_sum, s_max, hits, rounds) in res.items():
        print('%-40s %6d %10d %18.3f %18.3f' % (k, NW, rounds, hits / 32 / s_sum, hits / 32 / s_max))
