#!/bin/bash
# Round check on a B200 box (run under gpurun): GPU parity tests, both bench arms, launch list + full ncu capture of
# one timed step.  Everything lands in gpurun_out/.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
SNAP=tests/golden/c2_settled_n100000_seed100_it60.npz
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
if [ ! -f $SNAP ]; then
  python bench.py --steps 5 --warmup 3 --no-cpu --save-settled $SNAP > /dev/null 2> gpurun_out/settle_$TAG.err
  cp $SNAP gpurun_out/
fi
timeout 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv env B2C_GRAPH=0 python bench.py --steps 3 --warmup 3 --no-cpu --profile-step > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_$TAG \
    env B2C_GRAPH=0 B2C_OVERLAP=0 python bench.py --steps 3 --warmup 3 --no-cpu --profile-step > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -12
