"""Dev tool: summarise an FG_CHOL_TRACE dump of k_chol_rs (per-unit %globaltimer stamps).
columns: slot phase sn r0 r1 nblk ncols leaf t_start t_lastready t_updates_done t_potrf_done t_solve_done (unused) t_flag"""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.int64)
for ph in (0, 1):
    m = a[a[:, 1] == ph]
    if not len(m):
        continue
    t0 = m[:, 8].min()
    T = (m[:, 8:15] - t0) / 1000.0          # us
    T[m[:, 8:15] == 0] = np.nan
    print('phase', 'AC'[ph], 'units', len(m), 'span %.1f us' % np.nanmax(T))
    d = {'load+fronts+updates (start->updates done)': T[:, 2] - T[:, 0], 'last ready -> updates done': T[:, 2] - T[:, 1],
         'potrf': T[:, 3] - T[:, 2], 'solve': T[:, 4] - T[:, 3], 'store + fence + flag': T[:, 6] - T[:, 4]}
    for k, v in d.items():
        print('   %-45s mean %7.2f  median %7.2f  p90 %7.2f us' % (k, np.nanmean(v), np.nanmedian(v), np.nanpercentile(v, 90)))
    head = m[:, 3] == 0                      # head units (diagonal block + first rows) against the further row blocks
    for nm, sel in (('head units', head), ('row blocks', ~head)):
        if sel.any():
            print('   %-12s n %5d  updates-done -> factor/diag-ready median %6.2f  -> solved %6.2f  -> flagged %6.2f us' % (
                nm, sel.sum(), np.nanmedian((T[:, 3] - T[:, 2])[sel]), np.nanmedian((T[:, 4] - T[:, 3])[sel]), np.nanmedian((T[:, 6] - T[:, 4])[sel])))
    # chain step: per leaf (phase A), time between consecutive supernodes' flags
    if ph == 0:
        for leaf in np.unique(m[:, 7])[:3]:
            ml = m[(m[:, 7] == leaf) & (m[:, 14] > 0)]
            done = {}
            for r in ml:
                done[r[2]] = max(done.get(r[2], 0), (r[14] - t0) / 1000.0)      # a supernode is done when its last block is
            fl = np.sort(np.array(list(done.values())))
            print('   leaf', leaf, 'supernodes', len(fl), 'done-to-done median %.2f us, total %.1f us' % (np.median(np.diff(fl)), fl[-1] - fl[0]))
