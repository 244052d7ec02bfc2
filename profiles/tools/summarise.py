"""Turns the round's ncu captures (gpurun_out/<round>_launches.csv, gpurun_out/<round>_full.ncu-rep; round tag from $ROUND, default r2) into the committed summaries:
profiles/<round>_launches.csv, _launch_summary.md, _ncu_full_summary.md, _traffic.json."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
R = os.environ.get('ROUND', 'r2')

def short(k):
    return k.split('(')[0].strip()

# ---- launch list
rows = list(csv.reader(open(os.path.join(G, R + '_launches.csv'))))
i0 = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[i0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
seq = []
for r in rows[i0 + 1:]:
    if len(r) < len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    ms = float(r[ix['Metric Value']]) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(r[ix['Metric Unit']], 1e-6)
    k = short(r[ix['Kernel Name']])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
    seq.append(k)
tot = sum(a[1] for a in agg.values())
shutil.copy(os.path.join(G, R + '_launches.csv'), os.path.join(P, R + '_launches.csv'))
with open(os.path.join(P, R + '_launch_summary.md'), 'w') as f:
    f.write('# Round %s launch list summary (C5, 1xB200, final kernels)\n\n' % R[1:])
    f.write('Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/%s_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (profiles/tools/profile_round.sh).\n' % R)
    f.write('Raw list: `profiles/' + R + '_launches.csv`.  Per-launch times under ncu are cold-cache and serialised: compare SHARES with the\nCUDA-event phase times `bench.py` prints (`phases_ms`, `roofline_all`).\n\n')
    f.write('| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n')
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write('| `%s` | %d | %.3f | %.3f | %.1f%% |\n' % (k, n, ms, ms / n, 100 * ms / tot))
    f.write('\nKernel launches in the list: %d (one graph upload with preintegration, %d `k_chol_rs` launches = two per factorisation).\n' % (len(seq), sum(v[0] for k, v in agg.items() if 'k_chol_rs' in k)))

# ---- full capture
out = subprocess.run(['ncu', '-i', os.path.join(G, R + '_full.ncu-rep'), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static']
seen = collections.OrderedDict()
traffic = {}
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    k = short(d['Kernel Name'])
    tag = k + (' (grid %s)' % d.get('launch__grid_size', '?'))
    if 'k_chol_rs' in k:       # two launches per factorisation, same grid: the leaf phase is the one that moves more L2 sectors
        n_seen = sum('k_chol_rs' in t for t in seen)
        if n_seen >= 2:
            continue
        tag = k + ' #%d (grid %s, %.0f M L2 sectors)' % (n_seen + 1, d.get('launch__grid_size', '?'), float(d['lts__t_sectors.sum']) / 1e6)
    if tag in seen:
        continue
    st = sorted([(float(v), h) for h, v in d.items() if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and v], reverse=True)[:5]
    seen[tag] = (d, u, st)
    mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
    rd = float(d['dram__bytes_read.sum']) * mult.get(u['dram__bytes_read.sum'], 1.0)
    wr = float(d['dram__bytes_write.sum']) * mult.get(u['dram__bytes_write.sum'], 1.0)
    traffic.setdefault(('k_schur_tiles' if 'k_schur_tiles' in k else k.replace('void ', '')), []).append(rd + wr)
with open(os.path.join(P, R + '_ncu_full_summary.md'), 'w') as f:
    f.write('# ncu --set full summary, round %s final kernels (C5, 1xB200)\n\n' % R[1:])
    f.write("`ncu --set full --clock-control none --import-source on -k 'regex:k_chol_rs|k_schur_tiles|k_backsolve_w|k_front_syrk|k_proj_obs|k_zmat|k_proj_pose|k_lm_backsub_obs|k_schur_rhs|k_preintegrate' -s 30 -c 16 python bench.py --steps 1 --warmup 1 --no-cpu-baseline` (profiles/tools/profile_round.sh)\n\n")
    f.write('One launch per kernel (values per launch; cold-cache, serialised).  `k_chol_rs` appears twice: the leaf phase (first) and the separator phase (second).  Interpretation: DESIGN.md section 6.\n\n')
    for tag, (d, u, st) in seen.items():
        f.write('## %s\n' % tag)
        for k in keys:
            if k in d:
                f.write('- %s: %s %s\n' % (k, d[k], u.get(k, '')))
        f.write('- top stalls (warps per issue): ' + ', '.join('%s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, h in st) + '\n\n')
json.dump({'source': 'profiles/' + R + '_ncu_full_summary.md (ncu --set full, C5, per launch; k_chol_rs = leaf phase + separator phase)',
           'dram_bytes_per_launch': {k: (sum(v) if 'k_chol_rs' in k else v[0]) for k, v in traffic.items()}},
          open(os.path.join(P, R + '_traffic.json'), 'w'), indent=1)
print(open(os.path.join(P, R + '_launch_summary.md')).read())
