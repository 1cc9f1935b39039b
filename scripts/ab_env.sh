# Dev tool (GPU): bench lines under different environment switches
# usage: bash scripts/ab_env.sh "<VAR=val ...>" "<configs>"     ('-' = no switch)
for v in $1; do
  for cfg in $2; do
    extra=""; if [ $cfg = C4 ]; then extra="--points 524288"; fi
    ( if [ "$v" != "-" ]; then export $v; fi; python bench.py --config $cfg --no-cpu-baseline --no-also $extra 2>/dev/null ) | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('$v $cfg', '%.3e'%l['value'], round(l['ms_per_step'],3), {k: round(x,3) for k,x in r['stage_ms'].items()})"
  done
done
