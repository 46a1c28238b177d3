import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__inst_executed.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_lsu.sum','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.sum','sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_uniform.sum','sm__inst_executed_pipe_cbu.sum','sm__inst_executed_pipe_adu.sum']
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr: print('  ',w, r[hdr.index(w)], units[hdr.index(w)])
    st=[(float(r[i]) if r[i] else 0,h) for i,h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    for v,h in sorted(st,reverse=True)[:7]: print('   stall',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),round(v,2))
