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
# gpurun_out/ travels back only up to 64 MiB: KEEP_REP=0 turns the report into the raw / source CSV pages on the box
if [ "${KEEP_REP:-1}" = "0" ]; then
  ncu -i gpurun_out/prof_${WL}_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_${WL}_$TAG.csv 2>/dev/null
  if [ -n "${SRC_KERNEL:-}" ]; then
    ncu -i gpurun_out/prof_${WL}_$TAG.ncu-rep --page source --print-source cuda,sass --csv --kernel-name "$SRC_KERNEL" --launch-count 1 > gpurun_out/src_${WL}_$TAG.csv 2>/dev/null
  fi
  rm -f gpurun_out/prof_${WL}_$TAG.ncu-rep
  ls -la gpurun_out/raw_${WL}_$TAG.csv
fi
