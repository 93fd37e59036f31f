#!/bin/bash
# runs ON THE GPU BOX: one `ncu --set full` capture of the kernels matching <regex> in one warm frame of <workload>,
# exported as raw / source CSV next to the report and the JIT cubins (for tools/ncu_regions.py, tools/sass_lines.py).
#   tools/gpu_ncu.sh <outdir> <workload> <kernel regex> [ENV=VAL ...]
set -u
O=gpurun_out/$1; W=$2; K=$3; shift 3
mkdir -p $O
env "$@" VB200_DUMP_CUBIN=$O/$W timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"$K" -s 9 -c 3 -o $O/prof_$W -f \
    python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > $O/prof_$W.log 2>&1
ncu -i $O/prof_$W.ncu-rep --page raw --csv > $O/raw_$W.csv 2>/dev/null
ncu -i $O/prof_$W.ncu-rep --page source --csv > $O/source_$W.csv 2>/dev/null
ls -la $O | head -30
