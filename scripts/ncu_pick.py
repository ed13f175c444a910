"""print selected metrics of an .ncu-rep (ncu -i ... --page raw --csv)"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sectors_op_read.sum', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
rows = list(csv.reader(subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
H, U = rows[0], rows[1]
for V in rows[2:]:
    print(V[H.index('Kernel Name')][:80])
    for i, h in enumerate(H):
        if h in WANT or 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(V[i] or 0) > 0.3 or (len(sys.argv) > 2 and sys.argv[2] in h):
            print('   %-90s %s %s' % (h, V[i], U[i]))
