#!/bin/bash
# ncu --set full capture of selected kernels in one timed step (run under gpurun, 1 GPU).
#   bash tools/gpu_ncu_full.sh TAG WORKLOAD 'regex:k_sweep|k_keys'
set -u
TAG=${1:-r02}; WL=${2:-c2}; KERN=${3:-regex:k_sweep}
mkdir -p gpurun_out
EXTRA=""
if [ "$WL" = "c2" ]; then EXTRA="--no-sharded"; fi
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$KERN" -o gpurun_out/prof_${WL}_$TAG -f \
    env B2C_GRAPH=0 B2C_OVERLAP=0 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu $EXTRA --profile-step > gpurun_out/ncu_full_${WL}_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full_${WL}_$TAG.log
ls -la gpurun_out/prof_${WL}_$TAG.ncu-rep
