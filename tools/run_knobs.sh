for spec in "" "12=0"; do
  QASR_DEBUG="$spec" python bench.py --steps 30 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('knobs [$spec]', d['ms_per_step'], d['e2e']['ms_per_step'])
"
done
