python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu 2>gpurun_out/t.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c2', d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'])"
done
