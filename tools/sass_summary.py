#!/usr/bin/env python
r"""SASS evidence of the tcgen05 / TMEM / TMA code paths in libsdab.so: per kernel, the counts of the
mnemonics /opt/skills/guides/B200_PROFILING.md names (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UTMASTG = TMA loads / stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier traffic).

    python tools/sass_summary.py > profiles/sass_libsdab.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = ['UTCHMMA.2CTA', 'UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'SYNCS', 'UTMAPF', 'MUFU', 'FFMA', 'STG', 'LDG', 'LDS', 'STS']


def main():
    lib = ROOT / 'sda_b200' / 'libsdab.so'
    out = subprocess.run(['cuobjdump', '-sass', str(lib)], capture_output=True, text=True).stdout
    cur, counts, total = None, collections.OrderedDict(), collections.Counter()

    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)

        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = cur.replace('sdab::(anonymous namespace)::', '').replace('void ', '')
            cur = re.sub(r'\((?!int\)|bool\)).*', '', cur)
            counts[cur] = collections.Counter()
            continue

        m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)

        if m and cur:
            op = m.group(1)
            total[cur] += 1

            for k in KEYS:
                if op == k or op.startswith(k + '.'):
                    if k == 'UTCHMMA' and op.startswith('UTCHMMA.2CTA'):
                        continue
                    counts[cur][k] += 1

    print(f'# cuobjdump -sass sda_b200/libsdab.so (sm_100a): instruction counts per kernel')
    print(f'{"kernel":64s} {"instr":>6s} ' + ' '.join(f'{k:>7s}' for k in KEYS[:9]))

    for name, c in counts.items():
        if not any(c[k] for k in KEYS[:9]):
            continue
        print(f'{name[:64]:64s} {total[name]:6d} ' + ' '.join(f'{c[k]:7d}' for k in KEYS[:9]))

    agg = collections.Counter()

    for c in counts.values():
        agg.update(c)

    print('# library totals: ' + ', '.join(f'{k} {agg[k]}' for k in KEYS[:9]))
    print(f'# kernels in the library: {len(counts)}; with tcgen05 / TMA instructions: '
          f'{sum(1 for c in counts.values() if any(c[k] for k in KEYS[:9]))}')


if __name__ == '__main__':
    main()
