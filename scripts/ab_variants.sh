# Dev tool (GPU): bench lines of several library variants (PDR_VARIANT builds) on several configurations
# usage: bash scripts/ab_variants.sh "<variant tags, '-' = product build>" "<configs>"
for v in $1; do
  lib=$PWD/pde_read_b200/libpderead_b200.so; if [ "$v" != "-" ]; then lib=$PWD/pde_read_b200/libpderead_b200_$v.so; fi
  for cfg in $2; do
    extra=""; if [ $cfg = C4 ]; then extra="--points 524288"; fi
    PDEREAD_LIB=$lib python bench.py --config $cfg --no-cpu-baseline --no-also $extra 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('$v $cfg', '%.3e'%l['value'], round(l['ms_per_step'],3), {k: round(x,3) for k,x in r['stage_ms'].items()})"
  done
done
