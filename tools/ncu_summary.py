"""Summarise an .ncu-rep (read on the CPU box): key throughput metrics + stall reasons -> JSON on stdout."""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
out = []
for d in rows[2:]:
    rec = {}
    for k in keep:
        if k in hdr:
            i = hdr.index(k); rec[k + (f" [{units[i]}]" if units[i] else "")] = d[i]
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
            try: stalls.append((float(d[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError: pass
    rec['stalls_per_issue'] = {n: round(v, 3) for v, n in sorted(stalls, reverse=True)[:8]}
    for i, h in enumerate(hdr):
        if 'pipe_fp64' in h and 'inst_executed' in h and h.endswith('.sum') or 'tensor' in h and 'pct' in h and 'dmma' in h:
            rec[h] = d[i]
    out.append(rec)
print(json.dumps({'report': rep, 'captures': out}, indent=1))
