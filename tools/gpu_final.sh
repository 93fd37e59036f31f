#!/bin/bash
# Developer aid, runs ON THE GPU BOX: the short version of tools/refresh_profiles.sh used after a late kernel
# change (GPU tests first, stop if they fail; then the C3/C4 bench lines and ncu captures, the other
# workloads as quick lines); with the argument c125: the C5 / C2 / C1 bench lines and ncu captures instead.
# Output goes to gpurun_out/refresh/ like the long version's.
set -u
O=gpurun_out/refresh
mkdir -p $O
if [ "${1:-}" = "c125" ]; then
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
  exit 0
fi
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
rc=$?
echo "pytest rc=$rc" >> $O/pytest.log
tail -6 $O/pytest.log
[ $rc -ne 0 ] && exit 1
VB200_SORT_RANK=1 timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > $O/quick_c4_ranksort.json 2> $O/quick_c4_ranksort.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python bench.py --workload c4 --steps 20 --warmup 5 > $O/bench_c4.json 2> $O/bench_c4.err
for w in c5 c6 c2; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/quick_$w.json 2> $O/quick_$w.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_c3.log 2>&1
VB200_DUMP_CUBIN=$O/c3 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"resolve|k_setup|k_vertex" -s 9 -c 3 -o $O/prof_c3 -f \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_c3.log 2>&1
VB200_DUMP_CUBIN=$O/c4 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"tile_ordered|k_sort|k_setup" -s 9 -c 3 \
    -o $O/prof_c4 -f python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_c4.log 2>&1
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/bench_*.json')+glob.glob(sys.argv[1]+'/quick_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],4),'val',round(d['value'],1),'phase',{k:round(v,4) for k,v in d.get('phase_ms',{}).items()},'e2e',d['e2e'].get('ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
