#!/usr/bin/env python
"""Developer aid: join an `ncu --page source --csv` SASS export with `nvdisasm -g` line info of the
same linked cubin, and rank CUDA source lines by executed warp instructions / stall samples.
usage: tools/sass_lines.py <ncu_sass.csv> <linked.cubin> <kernel_name> [top_n]
(kernel_name is matched as a substring of both the demangled ncu name and the mangled .text section)"""
import csv, re, subprocess, sys, os
csvp, cubin, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
addr2line, cur, insec = {}, None, False
for l in out.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m:
        insec = (kern in m.group(1))
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
allrows = list(csv.reader(open(csvp)))
# the export holds one block per profiled kernel: ["Kernel Name", name], header row, instruction rows
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"]
blocks = [(allrows[i][1], allrows[i:(starts[k + 1] if k + 1 < len(starts) else len(allrows))]) for k, i in enumerate(starts)]
rows = [b for n, b in blocks if kern in n][-1]
h = rows[1]
ai, ie, si, ti = h.index('Address'), h.index('Instructions Executed'), h.index('# Samples'), h.index('Thread Instructions Executed')
base, byline, tot, tots = None, {}, 0, 0
for r in rows[2:]:
    try:
        a, n, s, tn = int(r[ai], 16), int(r[ie]), int(r[si]), int(r[ti])
    except Exception:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base, (None, ''))[0]
    d = byline.setdefault(key, [0, 0, 0])
    d[0] += n; d[1] += s; d[2] += tn
    tot += n; tots += s
print("total warp instr", tot, "samples", tots, "lines", len(byline))
srcs = {}
def text(key):
    if not key: return ''
    f, ln = key
    for d in ("visor_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            return srcs[p][ln - 1].strip()[:95] if ln - 1 < len(srcs[p]) else ''
    return ''
for key, (n, s, tn) in sorted(byline.items(), key=lambda kv: -kv[1][0])[:topn]:
    print("%5.1f%% instr %5.1f%% stall thr/inst %4.1f  %-28s %s" % (100 * n / tot, 100 * s / max(tots, 1), tn / max(n, 1), key, text(key)))
