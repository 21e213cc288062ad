import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.per_cycle_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed.sum','sm__cycles_elapsed.max','smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_bytes.sum','l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum']
for r in rows[2:3]:
    print("kernel", r[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr: print(f"  {w:72s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
    stall=[h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
    vals=sorted(((float(r[hdr.index(h)].replace(',','')),h) for h in stall), reverse=True)[:7]
    for v,h in vals: print(f"  {h.replace('smsp__average_warp','').replace('_per_warp_active.pct',''):72s} {v:.2f}")
