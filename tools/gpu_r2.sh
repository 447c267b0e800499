#!/bin/bash
# One GPU call of round 2 (run under gpurun, 1 GPU): all GPU tests, both bench arms, ncu launch lists of one C2 / C5 step.
set -u
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 1500 python -m pytest tests -m gpu -q --durations=15 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu_$TAG.log
fi
if [ "${SKIP_TESTS:-0}" != "1" ]; then B2C_SORT=radix timeout 600 python -m pytest tests -m gpu -q -k "c2_bin or c5_spheres or c4_batched or c1_stack or partitioned_bin" > gpurun_out/pytest_gpu_radix_$TAG.log 2>&1; echo "pytest radix rc=$?"; tail -3 gpurun_out/pytest_gpu_radix_$TAG.log; fi
timeout 600 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
    print('C2 ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'd2h', d['e2e']['d2h_bytes_per_step'], 'roofline', round(d['roofline']['frac'],4), d['roofline']['ms'])
    print(d['stage_ms'])
    for k in ('c4','c5'):
        if k in d: print(k, d[k]['ms_per_step'], 'e2e', d[k]['e2e_ms_per_step'], d[k]['stage_ms_rank0'], d[k].get('check'))
    print('cpu', d.get('cpu_baseline'))
except Exception as e: print('no bench json', e)
PY
if [ "${SKIP_REF:-0}" != "1" ]; then
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_$TAG.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2_$TAG.csv \
    env B2C_GRAPH=0 B2C_OVERLAP=0 python bench.py --steps 3 --warmup 3 --no-cpu --no-sharded --profile-step > gpurun_out/ncu_c2_$TAG.log 2>&1; echo "ncu c2 rc=$?"
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c5_$TAG.csv \
    env B2C_GRAPH=0 B2C_OVERLAP=0 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu --profile-step > gpurun_out/ncu_c5_$TAG.log 2>&1; echo "ncu c5 rc=$?"
fi
ls -la gpurun_out | tail -12
