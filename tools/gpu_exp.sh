#!/bin/bash
# round-2 iteration script (runs on the GPU box): GPU tests, then bench runs named on the command line.
#   tools/gpu_exp.sh <outdir> [notest] -- "<tag> <workload> [ENV=VAL ...]" ...
set -u
O=gpurun_out/$1; shift
mkdir -p $O
if [ "${1:-}" != "notest" ]; then
  timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
  echo "pytest rc=$?" >> $O/pytest.log
  tail -25 $O/pytest.log
else
  shift
fi
[ "${1:-}" = "--" ] && shift
for spec in "$@"; do
  set -- $spec
  tag=$1; w=$2; shift 2
  env "$@" timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}_$tag.json 2> $O/bench_${w}_$tag.err
done
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],4),'phase',{k:round(v,4) for k,v in d['phase_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3), d.get('parity'), 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
