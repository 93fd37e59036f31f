#!/bin/bash
# round-2 iteration script (runs on the GPU box): GPU tests, then bench c3/c5/c2/c1 with the direct visibility
# path on (default), off, and at other size thresholds.  usage: tools/gpu_r2d.sh <outdir name> [notest]
set -u
O=gpurun_out/${1:-r2d}
mkdir -p $O
if [ "${2:-}" != "notest" ]; then
  timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
  echo "pytest rc=$?" >> $O/pytest.log
  tail -15 $O/pytest.log
fi
for opt in "" "direct_visibility=0" "direct_max_pixels=32" "direct_max_pixels=128" "direct_max_pixels=256"; do
  tag=$(echo "${opt:-default}" | tr '=' '_')
  for w in c3 c5 c2 c1; do
    VB200_OPTIONS="$opt" timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}_$tag.json 2> $O/bench_${w}_$tag.err
  done
done
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],4),'phase',{k:round(v,4) for k,v in d['phase_ms'].items()}, d.get('parity'), d.get('fragments'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
