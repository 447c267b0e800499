#!/bin/bash
# 8-GPU scaling points (run under gpurun --gpus 8): C2 weak, C4 strong, C5 strong.
set -u
N=${1:-8}; TAG=${2:-r01}
mkdir -p gpurun_out
run() {
  wl=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N \
      --workload $wl --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/mg_${wl}_n${N}_$TAG.json 2> gpurun_out/mg_${wl}_n${N}_$TAG.err
  echo "== $wl gpus=$N rc=$?"; tail -1 gpurun_out/mg_${wl}_n${N}_$TAG.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['pairs_per_step'], d['stage_ms'])
except Exception as e: print('no json', e)"
}
run c2
run c4
run c5 --bodies 1000000 --max-pairs 8388608
grep -h -i "error\|Traceback" gpurun_out/mg_*_n${N}_$TAG.err | head -5
