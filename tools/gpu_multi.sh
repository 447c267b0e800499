#!/bin/bash
# Multi-GPU check (run under gpurun --gpus N): partition tests, then C5 (one partitioned world, strong scaling),
# C4 (batched worlds split by world, strong scaling) and C2 (one world per GPU, weak scaling) at 1 and N GPUs.
set -u
N=${1:-2}
TAG=${2:-r01}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_partitioned.py -m gpu -x -q 2>&1 | tail -3
run() {  # workload gpus extra...
  wl=$1; g=$2; shift 2
  if [ "$g" = 1 ]; then
    python bench.py --gpus 1 --workload $wl --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/mg_${wl}_n1_$TAG.json 2> gpurun_out/mg_${wl}_n1_$TAG.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $g \
      --workload $wl --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/mg_${wl}_n${g}_$TAG.json 2> gpurun_out/mg_${wl}_n${g}_$TAG.err
  fi
  echo "== $wl gpus=$g rc=$?"; tail -1 gpurun_out/mg_${wl}_n${g}_$TAG.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['pairs_per_step'], d['stage_ms'])
except Exception as e: print('no json', e)"
}
run c5 1 --bodies 1000000 --max-pairs 8388608
run c5 $N --bodies 1000000 --max-pairs 8388608
run c4 1
run c4 $N
run c2 $N
grep -h -i "error\|Traceback" gpurun_out/mg_*_$TAG.err | head
