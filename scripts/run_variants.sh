for v in "" _j1r64 _j1r80 _j1r64p4; do
  echo "== variant [$v]"
  PDEREAD_LIB=$PWD/pde_read_b200/libpderead_b200$v.so timeout 120 python scripts/step_times.py C2 C1 2>&1 | grep '"tc": 0' | cut -c1-220
done
