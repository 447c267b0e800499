#!/bin/bash
# N-GPU bench line (run under gpurun --gpus N): the default line = C2 replicas + the sharded c4 / c5 objects with the
# union == single-GPU check inside the NCCL run.
set -u
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -$N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N \
    --steps 20 --warmup 5 ${BENCH_ARGS:-} > gpurun_out/mg_n${N}_$TAG.json 2> gpurun_out/mg_n${N}_$TAG.err
echo "rc=$?"; tail -5 gpurun_out/mg_n${N}_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/mg_n${N}_$TAG.json').read().strip().splitlines()[-1])
    print('C2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
    for k in ('c4','c5'):
        if k in d: print(k, 'ms', d[k]['ms_per_step'], 'e2e', d[k]['e2e_ms_per_step'], d[k]['stage_ms_rank0'], d[k].get('check'), d[k]['collective'])
except Exception as e: print('no json', e)
PY
if [ "${WITH_REF:-0}" = "1" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus $N \
    --steps 3 --warmup 1 > gpurun_out/mg_ref_n${N}_$TAG.json 2> gpurun_out/mg_ref_n${N}_$TAG.err; echo "ref rc=$?"; cat gpurun_out/mg_ref_n${N}_$TAG.json
fi
