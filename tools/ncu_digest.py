"""Digest of an exported ncu report: python tools/ncu_digest.py <raw.csv> [<source.csv>]  (CSV from `ncu -i X.ncu-rep
--page raw --csv` / `--page source --csv`).  Prints the metrics the roofline discussion uses, per captured launch, and
the SASS opcodes / instructions that collect the most stall samples."""
import collections
import csv
import json
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active']


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = {}
    names = [d[hdr.index('Kernel Name')][:60] for d in data]
    out['kernels'] = names
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out[w + ' [' + units[i] + ']'] = [d[i] for d in data]
    return out


def source(path, top=14):
    rows = list(csv.reader(open(path)))
    kern, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {'name': r[1], 'rows': []}
            kern.append(cur)
        elif r and r[0] == "Address":
            cur['hdr'] = r
        elif cur is not None and len(r) > 5:
            cur['rows'].append(r)
    res = []
    for k in kern:
        h = k['hdr']
        ia, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
        tot = sum(int(r[isamp]) for r in k['rows']) or 1
        totex = sum(int(r[iex]) for r in k['rows']) or 1
        byop = collections.Counter()
        for r in k['rows']:
            parts = r[ia].split()
            op = parts[1] if parts[0].startswith('@') else parts[0]
            byop[op] += int(r[isamp])
        hot = sorted(k['rows'], key=lambda r: -int(r[isamp]))[:top]
        res.append({'kernel': k['name'][:70], 'samples': tot, 'warp_instructions': totex,
                    'stall_samples_by_opcode_pct': {op: round(100 * c / tot, 1) for op, c in byop.most_common(10)},
                    'hottest': [[r[ia][:70], round(100 * int(r[isamp]) / tot, 1)] for r in hot]})
    return res


if __name__ == "__main__":
    print(json.dumps(raw(sys.argv[1]), indent=1))
    if len(sys.argv) > 2:
        print(json.dumps(source(sys.argv[2])[:1], indent=1))
