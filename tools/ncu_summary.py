"""Print the metrics we track from an ncu report:  python tools/ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct']
stall = [k for k in h if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('-' * 60)
    for k in keys:
        if k in h:
            print('%-66s %s %s' % (k, r[h.index(k)][:80], rows[1][h.index(k)]))
    st = sorted(((float(r[h.index(k)] or 0), k) for k in stall), reverse=True)[:8]
    for v, k in st:
        print('   stall %-50s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
