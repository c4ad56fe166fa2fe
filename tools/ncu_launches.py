#!/usr/bin/env python
r"""Summarises an `ncu --csv --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...,dram__bytes_*`
launch list (one row per kernel launch and metric) per kernel and problem size.

    python tools/ncu_launches.py gpurun_out/launches.csv > profiles/<name>.txt

Launches of one kernel are split into clusters of similar DRAM traffic (the same template instance serves
several layer shapes).  Per-launch times under ncu are cold-cache and serialised: read shares, not step times."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]

    r = csv.reader(lines)
    hdr = next(r)
    d = collections.defaultdict(dict)

    for row in r:
        rec = dict(zip(hdr, row))
        i = int(rec['ID'])
        d[i]['name'] = rec['Kernel Name']
        d[i][rec['Metric Name']] = float(rec['Metric Value'].replace(',', ''))

    T = 'gpu__time_duration.sum'
    TP = 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'
    total = sum(v.get(T, 0) for v in d.values()) / 1e6
    groups = collections.defaultdict(list)

    for v in d.values():
        name = v['name'].split('(')[0].replace('void ', '').replace('sdab::<unnamed>::', '')[:48]
        dram = (v.get('dram__bytes_read.sum', 0) + v.get('dram__bytes_write.sum', 0)) / 1e9
        groups[(name, round(dram, 1) if 'conv_umma' in name else 0)].append((v.get(T, 0) / 1e6, dram, v.get(TP, 0)))

    print(f'# {path}: {len(d)} launches, {total:.1f} ms under ncu (cold-cache, serialised)')
    print(f'{"kernel":50s} {"dram GB":>8s} {"n":>5s} {"avg ms":>8s} {"total ms":>9s} {"share":>6s} {"tensor %":>8s}')
    rows = sorted(groups.items(), key=lambda kv: -sum(x[0] for x in kv[1]))

    for (name, _), l in rows:
        ms = sum(x[0] for x in l)
        if ms / total < 0.001:
            continue
        print(f'{name:50s} {sum(x[1] for x in l) / len(l):8.2f} {len(l):5d} {ms / len(l):8.3f} {ms:9.2f} {100 * ms / total:5.1f}% '
              f'{sum(x[2] * x[0] for x in l) / max(ms, 1e-9):8.1f}')

    conv = sum(x[0] for (n, _), l in groups.items() if 'conv_umma' in n for x in l)
    tp = sum(x[0] * x[2] for (n, _), l in groups.items() if 'conv_umma' in n for x in l) / max(conv, 1e-9)
    print(f'# convolution kernels: {100 * conv / total:.1f}% of the GPU time, time-weighted tensor-pipe active {tp:.1f}%')


if __name__ == '__main__':
    main(sys.argv[1])
