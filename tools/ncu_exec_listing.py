#!/usr/bin/env python
"""Developer aid: SASS of one kernel annotated with executed warp-instruction counts and source lines, plus a
summary of consecutive instructions executed about equally often (= the loops of the kernel).
usage: tools/ncu_exec_listing.py <ncu --page source --csv export> <cubin> <kernel substring> [listing.txt]"""
import csv, re, subprocess, sys
csvp, cubin, kern = sys.argv[1:4]
out = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
cur, rows, insec = None, [], False
for l in out.splitlines():
    if l.startswith('\t.section'):
        insec = ('.text.' in l) and (kern in l)
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        rows.append((int(m.group(1), 16), cur, m.group(2).strip()))
r = list(csv.reader(open(csvp)))
starts = [i for i, x in enumerate(r) if x and x[0] == "Kernel Name"]
blocks = [(r[i][1], r[i:(starts[k + 1] if k + 1 < len(starts) else len(r))]) for k, i in enumerate(starts)]
rr = [b for n, b in blocks if kern in n][-1]
h = rr[1]
ai, ie, si = h.index('Address'), h.index('Instructions Executed'), h.index('# Samples')
base, ex, st = None, {}, {}
for x in rr[2:]:
    try:
        a, n, s = int(x[ai], 16), int(x[ie]), int(x[si])
    except Exception:
        continue
    if base is None:
        base = a
    ex[a - base], st[a - base] = n, s
if len(sys.argv) > 4:
    with open(sys.argv[4], "w") as f:
        for a, k, ins in rows:
            f.write(f"{a:05x} {ex.get(a, 0):9d} {st.get(a, 0):5d} {(k[0][:16] if k else '?'):16s}:{(k[1] if k else 0):4d}  {ins[:100]}\n")
groups = []
for a, k, ins in rows:
    n = ex.get(a, 0)
    ln = k[1] if k and k[0].endswith('.cu') else None
    if groups and abs(groups[-1]['avg'] - n) <= max(0.08 * groups[-1]['avg'], 50):
        g = groups[-1]
        g['cnt'] += 1; g['sum'] += n; g['avg'] = g['sum'] / g['cnt']; g['st'] += st.get(a, 0)
        if ln: g['lines'].append(ln)
    else:
        groups.append({'start': a, 'cnt': 1, 'sum': n, 'avg': n, 'st': st.get(a, 0), 'lines': [ln] if ln else []})
tot = sum(g['sum'] for g in groups); tots = sum(g['st'] for g in groups)
print(f"total {tot / 1e6:.2f} M warp instructions, {tots} stall samples")
for g in groups:
    if g['sum'] > tot * 0.004:
        ls = g['lines']
        print(f"{g['start']:05x} n={g['cnt']:4d} x {g['avg']:9.0f} = {g['sum'] / 1e6:6.2f} M ({100 * g['sum'] / tot:4.1f}%)  stall {100 * g['st'] / max(tots, 1):4.1f}%  lines {min(ls) if ls else ''}-{max(ls) if ls else ''}")
