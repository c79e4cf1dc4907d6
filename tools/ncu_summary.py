#!/usr/bin/env python
"""Summarise ncu output for profiles/: `launches` = per-kernel share of a gpu__time_duration launch list (csv),
`raw` = key metrics per captured launch from `ncu -i rep --page raw --csv`."""
import collections, csv, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]
    kn, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
            continue
        name = r[kn].split('(')[0].replace('void ', '').replace('mccnn::<unnamed>::', '')
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    print(f'# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.2f} ms (ncu-serialised, cold cache: compare shares)')
    print('share%  launches  avg_ms  kernel')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{v[1] / tot * 100:6.2f}  {v[0]:5d}  {v[1] / v[0] / 1e6:9.3f}  {k[:100]}')


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = collections.Counter()
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        seen[name] += 1
        if seen[name] > 2:
            continue
        print('----', name[:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'  {k:80s} {r[i]:>16s} {units[i]}')
        for i, h in enumerate(hdr):
            if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct') and float(r[i] or 0) > 5:
                print(f'  {h:80s} {r[i]:>16s} %')


if __name__ == '__main__':
    {'launches': launches, 'raw': raw}[sys.argv[1]](sys.argv[2])
