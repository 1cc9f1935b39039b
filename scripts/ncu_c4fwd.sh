# Dev tool (GPU): ncu --set full of the C4 forward-U kernel for library variants
for v in $1; do
  lib=$PWD/pde_read_b200/libpderead_b200.so; if [ "$v" != "-" ]; then lib=$PWD/pde_read_b200/libpderead_b200_$v.so; fi
  PDEREAD_LIB=$lib ncu --set full --import-source on --clock-control none --kernel-name regex:mlp_jet_fwd --launch-skip 4 --launch-count 1 -f -o gpurun_out/c4fwd_$v python bench.py --config C4 --points 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/c4fwd_$v.log 2>&1
done
ls -la gpurun_out/c4fwd_*
