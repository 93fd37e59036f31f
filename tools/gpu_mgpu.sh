#!/bin/bash
# runs ON THE GPU BOX with N GPUs (gpurun --gpus N -- 'bash tools/gpu_mgpu.sh <outdir> N [test]'): the world-2 tests of
# the product's multi-GPU layer, then the bench at N ranks (fused exchange and the NCCL all-gather baseline)
set -u
O=gpurun_out/$1; N=$2
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
if [ "${3:-}" = "test" ]; then
  timeout 900 python -m pytest tests/test_mgpu.py -m gpu -x -q > $O/pytest_mgpu.log 2>&1
  echo "pytest rc=$?" >> $O/pytest_mgpu.log
  tail -5 $O/pytest_mgpu.log
fi
for n in $(seq 1 $N); do
  case $n in 1|2|4|8) ;; *) continue;; esac
  if [ $n = 1 ]; then
    timeout 900 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c5_n1.json 2> $O/bench_c5_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
        bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c5_n$n.json 2> $O/bench_c5_n$n.err
  fi
done
if [ $N -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29599 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --exchange allgather > $O/bench_c5_n${N}_allgather.json 2> $O/bench_c5_n${N}_allgather.err
fi
python - $O <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/bench_*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],4),'phase',{k:round(v,4) for k,v in d['phase_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3), d.get('parity'), d['config'].get('parallelism','')[:60])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
