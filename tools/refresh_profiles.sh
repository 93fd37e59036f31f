#!/bin/bash
# Developer aid, runs ON THE GPU BOX (gpurun -- 'bash tools/refresh_profiles.sh [notest]'): regenerates the raw
# material of profiles/ into gpurun_out/refresh/. Afterwards run tools/refresh_profiles.py <tag> here.
set -u
O=gpurun_out/refresh
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
if [ "${1:-}" != "notest" ]; then
  timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
  echo "pytest rc=$?" >> $O/pytest.log
  tail -4 $O/pytest.log
fi
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
for w in c1 c2 c4 c5 c6; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err
done
for w in c3 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_$w.log 2>&1
done
# full captures: every kernel of one warm frame of C3 / C5 / C2 (resolve path), the ordered path on C4
for w in c3 c5 c2; do
  VB200_DUMP_CUBIN=$O/$w timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:"resolve|k_setup|k_vertex" -s 9 -c 3 -o $O/prof_$w -f \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_$w.log 2>&1
done
VB200_DUMP_CUBIN=$O/c4 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"tile_ordered|k_sort|k_setup" -s 9 -c 3 \
    -o $O/prof_c4 -f python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_c4.log 2>&1
ls -la $O
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],4),'val',round(d['value'],1),'phase',{k:round(v,4) for k,v in d.get('phase_ms',{}).items()},'e2e',d['e2e'].get('ms_per_step'), d.get('parity'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
