#!/usr/bin/env python3
"""Summarises ncu output brought back in gpurun_out/ into small tracked files under profiles/.
   python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
   python scripts/summarize_ncu.py full gpurun_out/prof_x.ncu-rep profiles/r1_x.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src, errors='ignore') if l.startswith('"')))
    hdr = rows[0]
    ik, iv, im = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv or r[im] != 'gpu__time_duration.sum':
            continue
        name = r[ik].split('(')[0].replace('<unnamed>::', '')
        agg[name][0] += 1
        agg[name][1] += float(r[iv].replace(',', ''))
    unit = rows[1][hdr.index('Metric Unit')] if len(rows) > 1 else 'ns'
    tot = sum(v[1] for v in agg.values())
    with open(dst, 'w') as f:
        f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n')
        f.write('source: `%s` (%d launches, %s total %s). Per-launch times under ncu are cold-cache and serialised: '
                'compare SHARES, not absolutes.\n\n' % (src, sum(v[0] for v in agg.values()), '%.3f' % tot, unit))
        f.write('| kernel | launches | total (%s) | share |\n|---|---:|---:|---:|\n' % unit)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.1f | %.1f%% |\n' % (k, v[0], v[1], 100 * v[1] / tot))
    print(open(dst).read())


KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']


def full(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, 'w') as f:
        f.write('# ncu --set full summary\n\nsource: `%s` (read with `ncu -i ... --page raw --csv`)\n\n' % src)
        for r in rows[2:]:
            f.write('## %s\n\n| metric | value | unit |\n|---|---:|---|\n' % r[idx['Kernel Name']].replace('<unnamed>::', ''))
            for k in KEYS:
                if k in idx:
                    f.write('| %s | %s | %s |\n' % (k, r[idx[k]], units[idx[k]]))
            f.write('\n')
    print(open(dst).read()[:3000])


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
