"""Compact summary of an .ncu-rep (first profiled launch unless --all)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'sm__cycles_active.avg', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.avg.per_cycle_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__pcsamp_warps_issue_stalled', 'smsp__average_warps_issue_stalled', 'sm__pipe_tensor_cycles_active']
for r in rows[2:3] if '--all' not in sys.argv else rows[2:]:
    print('-' * 100)
    stalls = []
    for h, u, v in zip(hdr, units, r):
        if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued'):
            try: stalls.append((float(v), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError: pass
        elif any(h == k or (k in h and k in ('sm__pipe_tensor_cycles_active',)) for k in KEYS):
            print(f'{h:85s} {u:10s} {v[:70]}')
    tot = sum(s for s, _ in stalls) or 1
    print('stall samples:', ', '.join(f'{n} {100*s/tot:.0f}%' for s, n in sorted(stalls, reverse=True)[:8]))
