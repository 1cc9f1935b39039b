# Dev tool (GPU): per-kernel step times of the default library and of variant builds (PDR_VARIANT=<tag>):
#   bash scripts/run_variants.sh "" _tag1 _tag2
for v in "${@:-}"; do
  echo "== variant [$v]"
  PDEREAD_LIB=$PWD/pde_read_b200/libpderead_b200$v.so timeout 120 python scripts/step_times.py C2 C3 C4 C1 2>&1 | grep '"tc": 0' | cut -c1-260
done
