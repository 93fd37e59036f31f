#!/bin/bash
# round-2 iteration script (runs on the GPU box): tests, bench, ncu of the C3 tile kernel
# usage: tools/gpu_r2b.sh <outdir name> [variant ...]   (variants: visor_b200/build/scaffold_<v>.ptx)
set -u
O=gpurun_out/${1:-r2b}
shift
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
for w in c3 c2 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
done
for v in "$@"; do
  for w in c3 c5; do
    VB200_SCAFFOLD_PTX=visor_b200/build/scaffold_$v.ptx timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}_$v.json 2> $O/bench_${w}_$v.err
  done
done
VB200_DUMP_CUBIN=$O/c3 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"resolve|k_setup|k_vertex" -s 30 -c 3 -o $O/prof_c3 -f \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_c3.log 2>&1
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],4),'val',round(d['value'],1),'phase',{k:round(v,4) for k,v in d['phase_ms'].items()},'e2e_ms',round(d['e2e']['ms_per_step'],3), d.get('parity'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
