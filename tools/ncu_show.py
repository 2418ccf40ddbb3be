"""Prints the interesting raw metrics of every kernel in an .ncu-rep (read here, without a GPU).   python tools/ncu_show.py <rep> [regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
        'sm__throughput.avg.pct', 'gpu__dram_throughput.avg.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct',
        'sm__inst_executed_pipe_xu', 'smsp__inst_executed_pipe_xu', 'sm__pipe_xu', 'smsp__inst_executed_pipe_lsu',
        'sm__cycles_active.avg', 'smsp__cycles_active.avg']
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if pat and not pat.search(name):
        continue
    print('====', name[:100])
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(w) for w in WANT) and v not in ('', '0'):
            if 'stalled' in h and not h.endswith('per_issue_active.ratio'):
                continue
            if 'stalled' in h and float(v.replace(',', '')) < 0.3:
                continue
            print(f'  {h:88s} {v} {u}')
