import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(out.splitlines()))
hdr=r[0]; units=r[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','launch__grid_size','launch__block_size','launch__registers_per_thread','sm__inst_executed.sum','smsp__inst_executed.avg.per_cycle_active','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_lsu.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','sm__cycles_elapsed.max','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__cycles_active.avg']
stall=[h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
for row in r[2:]:
    print('-----')
    for w in want+stall:
        for i,h in enumerate(hdr):
            if h==w:
                v=row[i]
                if w in stall:
                    try:
                        if float(v)<3: continue
                    except: pass
                print('%-82s %s %s'%(h,v[:70],units[i]))
