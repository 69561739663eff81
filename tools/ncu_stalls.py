"""Summarise an .ncu-rep: headline metrics per kernel + stall reasons aggregated by CUDA source line.
   python tools/ncu_stalls.py report.ncu-rep [kernel-index] [top-n]"""
import csv, subprocess, sys, io, collections, re
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'smsp__issue_active.avg.pct', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit']
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want) and not re.search(r'\.(min|max)\b|peak_sustained$|per_second', h):
        print(f"{h:75s}", [r[i][:12] for r in rows[2:]])
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        vals = [r[i] for r in rows[2:]]
        if any(float(v or 0) > 0.3 for v in vals):
            print(f"{h.replace('smsp__average_warps_issue_stalled_', 'stall '):75s}", [v[:6] for v in vals])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
print(src[:0])
