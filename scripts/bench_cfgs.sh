# Dev tool (GPU): GPU parity tests, then one bench line per configuration with the per-kernel stage times.
# usage: bash scripts/bench_cfgs.sh <tag> [configs...]
tag=${1:-run}; shift
cfgs=${@:-C2 C3 C4 C5 C1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
for cfg in $cfgs; do
  extra=""; if [ $cfg = C4 ]; then extra="--points 524288"; fi
  python bench.py --config $cfg --no-cpu-baseline $extra 2>gpurun_out/${tag}_${cfg}.err | tee gpurun_out/${tag}_${cfg}.json | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('$cfg', '%.3e'%l['value'], round(l['ms_per_step'],3), {k: round(v,3) for k,v in r['stage_ms'].items()}, 'frac', round(r['frac'],3), 'whole', round(r['whole_step']['frac'],3))"
done
