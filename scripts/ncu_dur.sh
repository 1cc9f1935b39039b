# Dev tool (GPU): per-launch durations (ncu launch list) of one bench configuration for several library variants
# usage: bash scripts/ncu_dur.sh "<variant tags, '-' = product build>" <config> "<extra bench args>"
for v in $1; do
  lib=$PWD/pde_read_b200/libpderead_b200.so; if [ "$v" != "-" ]; then lib=$PWD/pde_read_b200/libpderead_b200_$v.so; fi
  PDEREAD_LIB=$lib ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,launch__registers_per_thread,sm__warps_active.avg.per_cycle_active --clock-control none --launch-skip 40 -c 30 --csv --log-file gpurun_out/dur_${v}_$2.csv python bench.py --config $2 --steps 2 --warmup 2 --no-cpu-baseline --no-also $3 > /dev/null 2>&1
  echo "== $v $2"; python scripts/ncu_launch_table.py gpurun_out/dur_${v}_$2.csv | awk '{print $2, $3, $4, $5, $6, $7, $8, $9, $10, $11}' | sort | uniq -c | sort -rn | head -12
done
