#!/bin/bash
# A/B of the halo exchange (peer-to-peer stores vs all-gather) on the partitioned C5 world (run under gpurun --gpus N).
set -u
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
for mode in ${MODES:-p2p fused nccl p2p fused nccl}; do
  hm=$mode; pm=1; if [ $mode = fused ]; then hm=p2p; pm=0; fi
  B2C_HALO=$hm B2C_HALO_PUSH=$pm timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N \
      --workload c5 --steps 30 --warmup 5 --no-cpu ${BENCH_ARGS:-} > gpurun_out/halo_${mode}_n${N}_$TAG.json 2> gpurun_out/halo_${mode}_n${N}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/halo_${mode}_n${N}_$TAG.json').read().strip().splitlines()[-1])
    s=d['stage_ms']
    print('$mode n=$N: step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), 'aabb+halo', s['aabb'], 'carry+migrate', s['unpack_carry'], 'sum', round(sum(s.values()),4))
except Exception as e: print('$mode failed', e); print(open('gpurun_out/halo_${mode}_n${N}_$TAG.err').read()[-1500:])
PY
done
