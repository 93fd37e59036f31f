#!/bin/bash
# variant sweep on the GPU box: tools/gpu_r2c.sh <outdir> "<variant>[:ENV=VAL,...]" ...   (variant "-" = built-in)
set -u
O=gpurun_out/${1:-sweep}
shift
mkdir -p $O
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo ${spec#*:} | tr ',' ' '); fi
  tag=$(echo $spec | tr ':,=' '___')
  for w in c3 c5; do
    if [ "$v" = "-" ]; then
      env $envs timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}_$tag.json 2> $O/bench_${w}_$tag.err
    else
      env $envs VB200_SCAFFOLD_PTX=visor_b200/build/scaffold_$v.ptx timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}_$tag.json 2> $O/bench_${w}_$tag.err
    fi
  done
done
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],4),'tiles',round(d['phase_ms']['tiles'],4), d.get('parity'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
