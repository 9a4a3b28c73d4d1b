for i in 1 2; do
  python bench.py --steps 30 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['plan'])
"
done
