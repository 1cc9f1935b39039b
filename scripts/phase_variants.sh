for v in "" _skipb _skipd _skipvjp; do
  for cfg in C2 C4; do
    extra=""; if [ $cfg = C4 ]; then extra="--points 524288"; fi
    PDEREAD_LIB=$PWD/pde_read_b200/libpderead_b200$v.so python bench.py --config $cfg --no-cpu-baseline --steps 10 $extra 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('$cfg [$v]', round(l['ms_per_step'],3), {k: round(v,3) for k,v in r['stage_ms'].items() if 'jet' in k})"
  done
done
