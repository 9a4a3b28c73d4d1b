for i in 1 2; do
for nof in 0 1; do
  QASR_NO_F32=$nof python bench.py --steps 30 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('no_f32=$nof', d['ms_per_step'], d['e2e']['ms_per_step'], d['plan'])
"
done
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
