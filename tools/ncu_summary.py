"""Key metrics per kernel from `ncu --page raw --csv`: python tools/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
idx = {h: i for i, h in enumerate(hdr)}
names = [r[idx['Kernel Name']].replace('void ', '').replace('dq::', '')[:60] for r in rows[2:]]
print(' ' * 52, names)
for w in want:
    if w in idx:
        vals = [r[idx[w]][:14] for r in rows[2:]]
        if w.startswith('smsp__average_warps') and all(float(v or 0) < 0.15 for v in vals):
            continue
        print(f"{w.replace('smsp__average_warps_issue_stalled_','stall:').replace('_per_issue_active.ratio','')[:52]:52s}", vals, units[idx[w]])
