import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
keys=['gpu__time_duration.sum','sm__cycles_elapsed.max','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_fmaheavy.sum','sm__inst_executed_pipe_fmalite.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_fp64.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','l1tex__lsu_writeback_active_mem_lg.sum','l1tex__lsu_writeback_active.sum','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum']
stall=[h for h in hdr if 'warp_issue_stalled' in h and 'per_warp_active' in h]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('====',d['Kernel Name'][:60])
    for k in keys:
        if k in d: print('  ',k,d[k])
    st=sorted([(float(d[h]),h) for h in stall if d[h] not in ('','n/a')],reverse=True)[:8]
    for v,h in st: print('   stall',h.replace('smsp__warp_issue_stalled_','').replace('_per_warp_active.pct',''),round(v,1))
