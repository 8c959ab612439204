#!/usr/bin/env python
"""
Turns ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python scripts/summarize_profile.py launches gpurun_out/X_launches.csv  > profiles/rNN_launches.md
  python scripts/summarize_profile.py full     gpurun_out/X.ncu-rep       > profiles/rNN_<kernel>_full.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tc.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'sm__cycles_active.avg']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split('(')[0][-90:]
        a = agg.setdefault((name, r[8], r[7]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1])
    total = sum(a[1] for a in agg.values())
    print(f'# ncu launch list: {path}\n')
    print('`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, not absolutes)\n')
    print('| kernel | grid | block | launches | avg us | total us | share |')
    print('|---|---|---|---:|---:|---:|---:|')
    for (name, grid, block), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{name}` | {grid} | {block} | {n} | {t / n / 1e3:.1f} | {t / 1e3:.1f} | {100 * t / total:.1f}% |')
    print(f'\ntotal {total / 1e3:.1f} us over {sum(a[0] for a in agg.values())} launches')


def full(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f'# ncu --set full: {path}\n')
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f'## {d.get("Kernel Name", "?")[:120]}  (launch id {d.get("ID")})\n')
        print('| metric | value | unit |')
        print('|---|---:|---|')
        for i, h in enumerate(hdr):
            if any(h == k or h.endswith('.' + k) or h == k + '.pct_of_peak_sustained_elapsed' for k in KEYS):
                print(f'| {h} | {r[i]} | {units[i]} |')
        print()


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
