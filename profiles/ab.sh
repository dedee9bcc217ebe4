#!/bin/bash
# A/B tuning runs in one gpurun call: each line of the here-doc is "label ENV=... ENV=..."; prints the
# step time and the top kernels of `bench.py --no-e2e` for each.   bash profiles/ab.sh <tag> <<< "..."
tag=${1:-ab}; shift
mkdir -p gpurun_out
while read -r label envs; do
  [ -z "$label" ] && continue
  env $envs timeout 600 python bench.py --steps ${STEPS:-2} --warmup ${WARMUP:-1} --no-cpu-baseline --no-e2e ${BENCH_ARGS} \
    > gpurun_out/${tag}_${label}.json 2> gpurun_out/${tag}_${label}.err
  python - "$label" gpurun_out/${tag}_${label}.json <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = l["kernels_ms_per_step"]
    print(f"{sys.argv[1]:14s} {l['ms_per_step']:8.2f} ms/step  frac {l['step_frac']:.3f} | " + " ".join(f"{a.replace('k_','')}={b:.1f}" for a, b in list(k.items())[:14]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[2].replace('.json', '.err')).read()[-400:])
PY
done
