import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
for vals in rows[2:]:
    d=dict(zip(hdr,vals))
    print('==',d.get('Kernel Name'))
    keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','launch__registers_per_thread','launch__waves_per_multiprocessor','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
    for k in keys:
        if k in d: print('  ',k,d[k])
    st=[(float(v),h) for h,v in d.items() if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    for v,h in sorted(st,reverse=True)[:7]: print('   stall',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),round(v,2))
    ops=[(h,v) for h,v in d.items() if 'sass_thread_inst_executed_op_d' in h and h.endswith('.sum')]
    for h,v in ops: print('  ',h,v)
