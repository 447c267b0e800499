python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c2', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches']/d['steps'], d['stage_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_tmp.csv python bench.py --steps 3 --warmup 3 --no-cpu --profile-step > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_tmp.csv')) if len(r)>10 and r[0].isdigit()]
tot=0
for r in rows:
    name=r[4].split('(')[0]; t=float(r[-1]); tot+=t
    if name.startswith('k_row') or name.startswith('k_carry'): print(f"{name[:50]:50s} {r[7]:>14s} {r[8]:>16s} {t/1000:8.2f} us")
print('total', tot/1000)
P
