#!/bin/bash
# Bench line + ncu launch list of one timed step for the non-headline configs (run under gpurun, 1 GPU).
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
run() {
  wl=$1; shift
  timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu "$@" > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
  echo "== $wl rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_${wl}_$TAG.json').read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['pairs_per_step'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['stage_ms'])"
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches_${wl}_$TAG.csv env B2C_GRAPH=0 B2C_OVERLAP=0 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --profile-step "$@" > /dev/null 2>&1
}
run c3 --bodies 10000
run c4
run c5 --bodies 1000000 --max-pairs 8388608
