#!/bin/bash
# Developer aid, runs ON THE GPU BOX (gpurun -- 'bash tools/refresh_profiles.sh'): regenerates the raw
# material of profiles/ into gpurun_out/refresh/. Afterwards run tools/refresh_profiles.py here.
set -u
O=gpurun_out/refresh
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
for w in c1 c2 c4 c5; do
  python bench.py --workload $w --steps 20 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_c3.log 2>&1
VB200_DUMP_CUBIN=$O/c3 ncu --set full --clock-control none --import-source on \
    -k regex:"resolve|k_setup|k_fill|k_scan|k_vertex|k_index_range" -s 24 -c 6 -o $O/prof_c3 -f \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_c3.log 2>&1
VB200_DUMP_CUBIN=$O/c4 ncu --set full --clock-control none --import-source on -k regex:"tile_ordered" -s 4 -c 1 \
    -o $O/prof_c4 -f python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_c4.log 2>&1
ls -la $O
