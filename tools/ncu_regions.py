#!/usr/bin/env python
"""Developer aid: executed warp instructions / stall samples of one kernel per SOURCE REGION.
usage: tools/ncu_regions.py <ncu --page source --csv export> <cubin> <kernel substring> <file> <line:name,...>
Regions are given as ascending 'first_line:name' pairs of <file>; instructions inlined from other files
count towards the region of the nearest preceding instruction of <file>."""
import collections, csv, re, subprocess, sys
csvp, cubin, kern, fname, spec = sys.argv[1:6]
bounds = [(int(a.split(':')[0]), a.split(':')[1]) for a in spec.split(',')]
def region(ln):
    name = bounds[0][1]
    for lo, nm in bounds:
        if ln >= lo:
            name = nm
    return name
out = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
addr2line, cur, insec = {}, None, False
for l in out.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m:
        insec = kern in m.group(1)
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csvp)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
blocks = [(rows[i][1], rows[i:(starts[k + 1] if k + 1 < len(starts) else len(rows))]) for k, i in enumerate(starts)]
rr = [b for n, b in blocks if kern in n][-1]
h = rr[1]
ai, ie, si, ti = h.index('Address'), h.index('Instructions Executed'), h.index('# Samples'), h.index('Thread Instructions Executed')
base, last = None, bounds[0][1]
reg = collections.OrderedDict((nm, [0, 0, 0]) for _, nm in bounds)
for r in rr[2:]:
    try:
        a, n, s, tn = int(r[ai], 16), int(r[ie]), int(r[si]), int(r[ti])
    except Exception:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base)
    if key and key[0] == fname and key[1] >= bounds[0][0]:
        last = region(key[1])
    d = reg[last]
    d[0] += n; d[1] += s; d[2] += tn
tot = sum(d[0] for d in reg.values()); tots = sum(d[1] for d in reg.values())
print(f"total warp instructions {tot / 1e6:.2f} M, stall samples {tots}")
for k, d in reg.items():
    print(f"  {k:26s} instr {d[0] / 1e6:7.2f} M {100 * d[0] / tot:5.1f}%   stall {100 * d[1] / max(tots, 1):5.1f}%   threads/instr {d[2] / max(d[0], 1):5.1f}")
