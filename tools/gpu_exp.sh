#!/bin/bash
# A/B of an environment knob on the C2 headline (run under gpurun, 1 GPU):  bash tools/gpu_exp.sh VAR "v0 v1 ..." [TAG]
set -u
VAR=$1; VALS=$2; TAG=${3:-exp}
mkdir -p gpurun_out
for v in $VALS; do
  for rep in 1 2; do
    env $VAR=$v timeout 300 python bench.py --no-cpu --no-sharded --steps 40 --warmup 5 > gpurun_out/exp_${TAG}_${v}_$rep.json 2> gpurun_out/exp_${TAG}_${v}_$rep.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/exp_${TAG}_${v}_$rep.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('$VAR=$v rep $rep: step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'gjk_mesh', d['stage_ms']['gjk_mesh'], 'epa', d['stage_ms']['epa_fold_count'], 'k_gjk', round(r.get('ms',0),4) if r.get('kernel')=='k_gjk' else None, 'frac', round(r['frac'],4))
except Exception as e: print('$VAR=$v failed', e)
PY
  done
done
