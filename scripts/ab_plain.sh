python -m pytest tests -m gpu -q 2>&1 | tail -4
for cfg in C2 C5 C4 C1; do
  for on in 1 0; do
    extra=""; if [ $cfg = C4 ]; then extra="--points 524288"; fi
    PDEREAD_PLAIN_MLP=$on python bench.py --config $cfg --no-cpu-baseline $extra 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('$cfg plain=$on', '%.3e'%l['value'], round(l['ms_per_step'],3), {k: round(v,3) for k,v in r['stage_ms'].items()})"
  done
done
