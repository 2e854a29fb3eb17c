"""Key metrics of every kernel in an .ncu-rep (ncu --set full): time, DRAM bytes, pipes, issue, stalls.  Usage: ncu_key.py rep [json_out]"""
import csv, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
hdr, units = r[0], r[1]
exact = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
         'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
         'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
         'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
         'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
         'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
         'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
         'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg']
res = []
for row in r[2:]:
    d = {}
    for i, h in enumerate(hdr):
        v = row[i]
        if h in exact:
            d[h] = v
        elif h.startswith('sm__inst_executed_pipe_') and h.endswith('.avg.pct_of_peak_sustained_active'):
            try:
                if float(v) >= 1: d[h] = v
            except ValueError: pass
        elif h.startswith('sm__pipe_') and h.endswith('cycles_active.avg.pct_of_peak_sustained_active'):
            try:
                if float(v) >= 1: d[h] = v
            except ValueError: pass
        elif 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
            try:
                if float(v) >= 3: d[h] = v
            except ValueError: pass
    res.append(d)
    print('-----')
    for k, v in d.items():
        print('%-90s %s' % (k, v[:80]))
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], 'w'), indent=1)
