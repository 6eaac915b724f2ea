"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum', 'smsp__inst_executed.sum']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('=== %s' % r[idx['Kernel Name']][:110])
    for w in want:
        if w in idx and r[idx[w]] not in ('', 'n/a'):
            print('   %-70s %s %s' % (w, r[idx[w]], units[idx[w]]))
    st = [(float(r[i].replace(',', '')), h.replace('smsp__pcsamp_warps_issue_stalled_', '')) for i, h in enumerate(hdr)
          if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('_not_issued') and r[i] not in ('', 'n/a')]
    tot = sum(v for v, _ in st) or 1
    print('   stalls: ' + ', '.join('%s %.0f%%' % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:6]))
