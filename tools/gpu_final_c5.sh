#!/bin/bash
# Developer aid, runs ON THE GPU BOX: C5 / C2 part of the short profile refresh (see tools/gpu_final.sh)
set -u
O=gpurun_out/refresh
mkdir -p $O
VB200_DUMP_CUBIN=$O/c5 timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:"resolve|k_setup|k_vertex" -s 9 -c 3 -o $O/prof_c5 -f \
    python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_c5.log 2>&1
timeout 300 python bench.py --workload c5 --steps 20 --warmup 5 > $O/bench_c5.json 2> $O/bench_c5.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c5.csv \
    python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_c5.log 2>&1
VB200_DUMP_CUBIN=$O/c2 timeout 200 ncu --set full --clock-control none --import-source on \
    -k regex:"resolve|k_setup|k_vertex" -s 9 -c 3 -o $O/prof_c2 -f \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_c2.log 2>&1
timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 200 python bench.py --workload c1 --steps 20 --warmup 5 > $O/bench_c1.json 2> $O/bench_c1.err
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/bench_c[125].json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],4),'val',round(d['value'],1),'phase',{k:round(v,4) for k,v in d.get('phase_ms',{}).items()},'e2e',d['e2e'].get('ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
