import csv, sys, subprocess
KEYS=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_active.avg.per_cycle_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum','l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.avg','l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct']
for rep in sys.argv[1:]:
    raw=subprocess.check_output(["ncu","-i",rep,"--page","raw","--csv"],text=True,stderr=subprocess.DEVNULL)
    rows=list(csv.reader(raw.splitlines())); H,U,R=rows[0],rows[1],rows[2]
    d={h:(r,u) for h,u,r in zip(H,U,R)}
    print("==",rep)
    for k in KEYS: print(f"  {k:90s} {d.get(k)}")
    for h in d:
        if 'issue_stalled' in h and 'per_issue_active' in h and float(d[h][0])>0.05: print(f"  {h[35:-23]:60s} {float(d[h][0]):.3f}")
