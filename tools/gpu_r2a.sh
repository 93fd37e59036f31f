#!/bin/bash
# round-2 iteration script (runs on the GPU box): tests, bench variants, ncu
set -u
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for mc in 5 4 0; do
  VB200_JIT_MINCTAS=$mc timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c3_mc$mc.json 2> $O/bench_c3_mc$mc.err
done
for w in c1 c2 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_c3.log 2>&1
VB200_DUMP_CUBIN=$O/c3 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"resolve|k_setup|k_vertex" -s 30 -c 3 -o $O/prof_c3 -f \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_c3.log 2>&1
ls -la $O
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],4),'val',round(d['value'],1),'phase',{k:round(v,4) for k,v in d['phase_ms'].items()},'e2e_ms',round(d['e2e']['ms_per_step'],3), d.get('parity'))
    except Exception as e:
        print(f, 'ERR', e)
PY
