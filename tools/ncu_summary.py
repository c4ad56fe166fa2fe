#!/usr/bin/env python
r"""Prints the headline metrics of every kernel in an ncu report (raw page) as a small table.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv, subprocess, sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('sm__cycles_elapsed.avg', 'sm cycles'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm throughput %'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram throughput %'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('l1tex__m_xbar2l1tex_read_bytes.sum', 'L2->SM bytes'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem wavefronts'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
]

rows = list(csv.reader(subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('kernel:', r[hdr.index('Kernel Name')][:150])
    for k, label in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f'  {label:24s} {r[i]:>16s} {units[i]}')
    print()
