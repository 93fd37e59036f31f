#!/usr/bin/env python
"""Developer aid: turn gpurun_out/refresh/ (written by tools/refresh_profiles.sh on the GPU box) into the
committed text/CSV/JSON evidence under profiles/.  usage: python tools/refresh_profiles.py [round_tag]"""
import csv, json, os, shutil, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))    # repo root
src = os.path.join(here, "gpurun_out", "refresh")
dst = os.path.join(here, "profiles")

def last_json(path):
    lines = [l for l in open(path).read().splitlines() if l.strip().startswith("{")]
    return json.loads(lines[-1])

for name, out in [("bench_c3", "bench_c3"), ("bench_reference_arm", "bench_reference_arm"), ("bench_c1", "bench_c1"),
                  ("bench_c2", "bench_c2"), ("bench_c4", "bench_c4"), ("bench_c5", "bench_c5_n1"),
                  ("bench_c6", "bench_c6_many_draws")]:
    p = os.path.join(src, name + ".json")
    if os.path.exists(p):
        json.dump(last_json(p), open(os.path.join(dst, f"{tag}_{out}.json"), "w"), indent=1)
        open(os.path.join(dst, f"{tag}_{out}.json"), "a").write("\n")

# launch lists: every kernel launch of the bench command with its ncu duration
for w in ("c3", "c5"):
    path = os.path.join(src, f"launches_{w}.csv")
    if not os.path.exists(path):
        continue
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, vi, ui, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
    with open(os.path.join(dst, f"{tag}_launches_{w}.csv"), "w") as f:
        wr = csv.writer(f)
        wr.writerow(["ID", "Kernel Name", "gpu__time_duration.sum", "unit"])
        for r in rows[1:]:
            wr.writerow([r[ii], r[ki], r[vi], r[ui]])

def summary(rep, out):
    if not os.path.exists(rep):
        return
    text = subprocess.run([sys.executable, os.path.join(here, "tools", "ncu_summary.py"), rep], capture_output=True,
                          text=True).stdout
    open(out, "w").write(text)

summary(os.path.join(src, "prof_c3.ncu-rep"), os.path.join(dst, f"{tag}_ncu_c3_kernels.txt"))
summary(os.path.join(src, "prof_c5.ncu-rep"), os.path.join(dst, f"{tag}_ncu_c5_kernels.txt"))
summary(os.path.join(src, "prof_c2.ncu-rep"), os.path.join(dst, f"{tag}_ncu_c2_kernels.txt"))
summary(os.path.join(src, "prof_c4.ncu-rep"), os.path.join(dst, f"{tag}_ncu_c4_kernels.txt"))

def hot_lines(rep, cubin, kernel, out):
    if not (os.path.exists(rep) and os.path.exists(cubin)):
        return
    srccsv = rep.replace(".ncu-rep", "_source.csv")
    with open(srccsv, "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=f, stderr=subprocess.DEVNULL)
    text = subprocess.run([sys.executable, os.path.join(here, "tools", "sass_lines.py"), srccsv, cubin, kernel, "45"],
                          capture_output=True, text=True).stdout
    open(out, "w").write(text)

hot_lines(os.path.join(src, "prof_c3.ncu-rep"), os.path.join(src, "c3.vb200_k_tile_resolve_min_first.cubin"),
          "resolve_min_first", os.path.join(dst, f"{tag}_ncu_c3_tile_resolve_hot_lines.txt"))
hot_lines(os.path.join(src, "prof_c5.ncu-rep"), os.path.join(src, "c5.vb200_k_tile_resolve_min_first.cubin"),
          "resolve_min_first", os.path.join(dst, f"{tag}_ncu_c5_tile_resolve_hot_lines.txt"))
hot_lines(os.path.join(src, "prof_c4.ncu-rep"), os.path.join(src, "c4.vb200_k_tile_ordered.cubin"),
          "tile_ordered", os.path.join(dst, f"{tag}_ncu_c4_tile_ordered_hot_lines.txt"))
def regions(rep, cubin, kernel, which, out):
    srccsv = rep.replace(".ncu-rep", "_source.csv")
    if not (os.path.exists(srccsv) and os.path.exists(cubin)):
        return
    text = subprocess.run([sys.executable, os.path.join(here, "tools", "ncu_regions.py"), srccsv, cubin, kernel, which],
                          capture_output=True, text=True).stdout
    open(out, "w").write(text)

regions(os.path.join(src, "prof_c3.ncu-rep"), os.path.join(src, "c3.vb200_k_tile_resolve_min_first.cubin"),
        "resolve_min_first", "resolve", os.path.join(dst, f"{tag}_ncu_c3_tile_resolve_regions.txt"))
regions(os.path.join(src, "prof_c5.ncu-rep"), os.path.join(src, "c5.vb200_k_tile_resolve_min_first.cubin"),
        "resolve_min_first", "resolve", os.path.join(dst, f"{tag}_ncu_c5_tile_resolve_regions.txt"))
regions(os.path.join(src, "prof_c4.ncu-rep"), os.path.join(src, "c4.vb200_k_tile_ordered.cubin"),
        "tile_ordered", "ordered", os.path.join(dst, f"{tag}_ncu_c4_tile_ordered_regions.txt"))
print("profiles/ refreshed:", sorted(os.listdir(dst)))
