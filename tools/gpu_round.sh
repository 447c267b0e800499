#!/bin/bash
# One GPU call of a round (run under gpurun): every GPU test (no -x: all failures in one go), both bench arms, the
# compound side measurement, and the ncu launch list of one timed step.  Everything lands in gpurun_out/.
set -u
TAG=${1:-r01k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
timeout 200 python tools/compound_bench.py 20000 30 > gpurun_out/compound_$TAG.json 2> gpurun_out/compound_$TAG.err; echo "compound rc=$?"
cat gpurun_out/compound_$TAG.json; tail -3 gpurun_out/compound_$TAG.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_$TAG.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv env B2C_GRAPH=0 python bench.py --steps 3 --warmup 3 --no-cpu --profile-step > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out | tail -12
