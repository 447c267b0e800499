#!/bin/bash
# A/B of an environment knob on the C5 config, one GPU (run under gpurun):  bash tools/gpu_exp_c5.sh VAR "v0 v1 ..." [TAG]
set -u
VAR=$1; VALS=$2; TAG=${3:-exp}
mkdir -p gpurun_out
for v in $VALS; do
    env $VAR=$v timeout 300 python bench.py --workload c5 --no-cpu --steps 30 --warmup 5 > gpurun_out/expc5_${TAG}_${v}.json 2> gpurun_out/expc5_${TAG}_${v}.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/expc5_${TAG}_${v}.json').read().strip().splitlines()[-1])
    s=d['stage_ms']
    print('$VAR=$v: step', round(d['ms_per_step'],4), 'gjk_mesh(=wait for k_sphere_sphere)', s['gjk_mesh'], 'epa_fold', s['epa_fold_count'], 'sweep', s['sweep'])
except Exception as e: print('$VAR=$v failed', e)
PY
done
