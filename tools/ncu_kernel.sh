#!/bin/bash
# one `ncu --set full` capture of one kernel of the bench step at 500x500x80 (NX=2000: at bench size); raw + source pages
# come back as CSV
#   [NX=2000] tools/ncu_kernel.sh <tag> <kernel-regex> [launch-skip]
tag=$1; kern=$2; skip=${3:-1}
ncu --set full --clock-control none --import-source on -k regex:"$kern" --launch-skip $skip --launch-count 1 -o /tmp/${tag} -f \
  python bench.py --steps 1 --warmup 1 --nx ${NX:-500} --ny ${NX:-500} --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_src.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page details > gpurun_out/${tag}_details.txt 2>/dev/null
