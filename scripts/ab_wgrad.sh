# Dev tool (GPU): step / kernel times with the weight-gradient product on the warp-level tensor path (1) or as FFMA2 tiles (0)
python -m pytest tests -m gpu -q 2>&1 | tail -4
for cfg in C2 C3 C4 C1; do
  for on in 1 0; do
    extra=""; if [ $cfg = C4 ]; then extra="--points 524288"; fi; if [ $cfg = C3 ]; then extra="--points 250000"; fi
    PDEREAD_WGRAD_MMA=$on python bench.py --config $cfg --no-cpu-baseline --steps 10 $extra 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('$cfg wgrad_mma=$on', '%.3e'%l['value'], round(l['ms_per_step'],3), {k: round(v,3) for k,v in r['stage_ms'].items()}, 'frac', round(r['frac'],3), 'whole', round(r['whole_step']['frac'],3))"
  done
done
