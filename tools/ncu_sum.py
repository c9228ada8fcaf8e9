import csv,sys,subprocess
rep=sys.argv[1]
out=open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; vals=rows[2]
d={h:(v,u) for h,u,v in zip(hdr,units,vals)}
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','launch__grid_size','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','sm__icc_request_hit_rate.pct','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum']
for k in keys:
    if k in d: print(k, d[k][1], d[k][0])
for h in hdr:
    if 'issue_stalled' in h and 'per_issue_active' in h or 'warps_issue_stalled' in h and h.endswith('.ratio') and 'per' in h:
        v=d[h][0]
        try:
            if float(v)>0.05: print(h.replace('smsp__average_warps_issue_stalled_','stall '), v)
        except: pass
