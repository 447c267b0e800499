python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in c2 c4; do
python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches']/d['steps'], d['stage_ms'])"
done
