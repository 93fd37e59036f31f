#!/usr/bin/env python
"""Developer aid: executed warp instructions, stall samples (with the top stall reasons) and shared-memory
wavefronts of one kernel per SOURCE REGION of scaffold.cu.
usage: tools/ncu_regions.py <ncu --page source --csv export> <cubin> <kernel substring> [resolve|ordered]
Regions are found by marker text in the current scaffold.cu (so the cubin must come from the same source);
instructions inlined from other files count towards the region of the nearest preceding instruction of
scaffold.cu; line 1 of scaffold.cu is the inlined shader (runtime.cpp pins its instructions there)."""
import collections, csv, os, re, subprocess, sys
csvp, cubin, kern = sys.argv[1:4]
which = sys.argv[4] if len(sys.argv) > 4 else "resolve"
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(here, "visor_b200", "csrc", "scaffold.cu")).read().splitlines()
MARKS = {
    "resolve": [("__device__ __forceinline__ uint32_t vb200_depth_key", "hit: depth / key / CAS helpers"),
                ("__device__ __forceinline__ void vb200_tile_resolve_body", "prologue, empty tiles"),
                ("for(uint32_t base = 0; base < n; base += RT)", "round: list, record and corner gathers, tile init"),
                ("const Vb200TriSetup su = vb200_unpack_setup(rq, ra, rb, rc);", "round: edge set-up, scan, staging"),
                ("const uint32_t ups = min(32u", "step: owner lookup"),
                ("---- 2. this lane's unit", "step: unit set-up"),
                ("const uint32_t wmax = __reduce_max_sync", "step: row walk"),
                ("const uint32_t colmask =", "step: coverage -> runs"),
                ("uint32_t ownerBase = 0;", "hit loop"),
                ("if(__any_sync(0xffffffffu, raggedA || raggedB))", "ragged rows"),
                ("---- phase B: shade the winner", "phase B")],
    "ordered": [("vb200_k_tile_ordered(const __grid_constant__", "prologue, region load"),
                ("auto shade_pass = ", "fragment pass"),
                ("for(uint32_t base = 0; base < n; base += 32u)", "triangle batch: load + set-up"),
                ("---- 2. surviving triangles in list order", "coverage -> ring"),
                ("for(int j = 0; j < 4; j++)\n", "write-back")],
}[which]
bounds = []
for text, name in MARKS:
    hits = [i + 1 for i, l in enumerate(src) if text.strip() in l]
    if hits:
        bounds.append((hits[0], name))
bounds.sort()
first = bounds[0][0]
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
col = {n: i for i, n in enumerate(h)}
stallcols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
names = ["shader (inlined)"] + [nm for _, nm in bounds]
reg = collections.OrderedDict((nm, collections.Counter()) for nm in names)
base, last = None, bounds[0][1]
def num(r, c):
    try:
        return int(r[col[c]] or 0)
    except Exception:
        return 0
for r in rr[2:]:
    try:
        a = int(r[col["Address"]], 16)
    except Exception:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base)
    where = last
    if key and key[0] == "scaffold.cu":
        if key[1] == 1:
            where = "shader (inlined)"
        elif key[1] >= first:
            last = where = region(key[1])
    d = reg[where]
    d["instr"] += num(r, "Instructions Executed")
    d["samples"] += num(r, "# Samples")
    d["threads"] += num(r, "Thread Instructions Executed")
    d["wf"] += num(r, "L1 Wavefronts Shared")
    d["gl"] += num(r, "L1 Tag Requests Global")
    for sc in stallcols:
        d[sc] += num(r, sc)
tot = sum(d["instr"] for d in reg.values()) or 1
tots = sum(d["samples"] for d in reg.values()) or 1
totw = sum(d["wf"] for d in reg.values()) or 1
print(f"{kern}: {tot / 1e6:.2f} M warp instructions, {tots} stall samples, {totw / 1e6:.2f} M shared-memory wavefronts")
print(f"{'region':46s} {'instr M':>8s} {'%':>6s} {'stall %':>8s} {'thr/ins':>8s} {'smem wf M':>10s} {'%':>6s} {'glob req M':>10s}  top stall reasons")
for k, d in reg.items():
    top = sorted(((sc, d[sc]) for sc in stallcols), key=lambda x: -x[1])[:3]
    print(f"{k:46s} {d['instr'] / 1e6:8.2f} {100 * d['instr'] / tot:6.1f} {100 * d['samples'] / tots:8.1f} "
          f"{d['threads'] / max(d['instr'], 1):8.1f} {d['wf'] / 1e6:10.2f} {100 * d['wf'] / totw:6.1f} {d['gl'] / 1e6:10.2f}  "
          + " ".join(f"{a[6:]}={100 * b / max(d['samples'], 1):.0f}%" for a, b in top))
