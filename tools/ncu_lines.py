#!/usr/bin/env python
r"""Maps the warp-stall samples of an ncu report (SASS level) to source lines of libsdab.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-substring] [top]

Needs the same libsdab.so that ran under ncu (built with -lineinfo)."""
import csv, re, subprocess, sys, tempfile, os
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else 'conv_umma_kernel'
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30

tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', str(ROOT / 'sda_b200' / 'libsdab.so')], cwd=tmp, capture_output=True)
dis = ''
for f in os.listdir(tmp):
    if f.endswith('.cubin'):
        out = subprocess.run(['nvdisasm', '-g', '-c', f], cwd=tmp, capture_output=True, text=True).stdout
        if kern in out:
            dis += out

src_csv = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src_csv.splitlines()))
name = rows[0][1] if rows and len(rows[0]) > 1 else ''
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
recs = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        recs.append((int(r[idx['Address']], 16), int(r[idx['# Samples']]), r))
    except ValueError:
        pass
base = min(a for a, _, _ in recs)

# pick the function section of the disassembly matching the profiled kernel (template args in the mangled name)
sections = re.split(r'\n(?=\s*\.section\s+\.text\.)', dis)
want = None
m = re.search(r'(\w+)<([^>]*)>', name)
mangled = None
if m:
    parts = []
    for a in m.group(2).split(','):
        a = a.strip()
        mm = re.match(r'\((\w+)\)(\d+)', a)
        ty, val = (mm.group(1), mm.group(2)) if mm else ('int', a)
        parts.append(('Lb' if ty == 'bool' else 'Li') + val + 'E')
    mangled = m.group(1) + 'I' + ''.join(parts) + 'E'
for sec in sections:
    head = sec[:600]
    if (mangled and mangled in head) or (mangled is None and kern in head):
        want = sec
        break
if want is None:
    want = dis
cur = None
off2line = {}
for l in want.splitlines():
    mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if mm:
        cur = (mm.group(1).split('/')[-1], int(mm.group(2)))
        continue
    mm = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if mm:
        off2line[int(mm.group(1), 16)] = (cur, mm.group(2).strip())

stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
byline, reasons = Counter(), {}
tot = 0
for a, s, r in recs:
    ln = off2line.get(a - base, (None, ''))[0]
    byline[ln] += s
    tot += s
    rc = reasons.setdefault(ln, Counter())
    for h in stall_cols:
        try:
            rc[h] += int(r[idx[h]] or 0)
        except ValueError:
            pass
print(name[:120], 'samples', tot)
for ln, s in byline.most_common(top):
    print(f'{100 * s / tot:5.1f}%  {ln}  {[k.replace("stall_", "") + ":" + str(v) for k, v in reasons[ln].most_common(2)]}')
