"""Per-kernel summary of an `ncu --set full ... --page raw --csv` dump: python tools/ncu_full_summary.py raw.csv "title" > out.csv"""
import csv
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum']
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = [hdr.index(w) for w in WANT if w in hdr]
w = csv.writer(sys.stdout)
if len(sys.argv) > 2:
    w.writerow(['# ' + sys.argv[2]])
w.writerow([hdr[i] for i in idx])
w.writerow([units[i] for i in idx])
for r in data:
    w.writerow([r[i][:60] if i == idx[0] else r[i] for i in idx])
