set -u
O=gpurun_out/r2e; mkdir -p $O
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_$tag.json 2> $O/bench_c3_$tag.err; }
run direct VB200_OPTIONS=
run noatom VB200_EXP_NOATOM=1
run pb2 VB200_SCAFFOLD_PTX=visor_b200/build/scaffold_pb2.ptx
run pb4 VB200_SCAFFOLD_PTX=visor_b200/build/scaffold_pb4.ptx
run pb2_256 VB200_SCAFFOLD_PTX=visor_b200/build/scaffold_pb2.ptx VB200_OPTIONS=direct_max_pixels=256
run binned_pb2 VB200_SCAFFOLD_PTX=visor_b200/build/scaffold_pb2.ptx VB200_OPTIONS=direct_visibility=0
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],4),'phase',{k:round(v,4) for k,v in d['phase_ms'].items()}, d.get('parity'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
