# Dev tool (GPU): ncu --set full of the U-network kernels from the bench command (C2 step, C5 extraction)
tag=${1:-ncu}
ncu --set full --import-source on --clock-control none --kernel-name regex:mlp_jet --launch-skip 8 --launch-count 2 -f -o gpurun_out/${tag}_U_c2 python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/${tag}_c2.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:mlp_jet_fwd --launch-skip 12 --launch-count 1 -f -o gpurun_out/${tag}_U_c5 python bench.py --config C5 --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/${tag}_c5.log 2>&1
ls -la gpurun_out/${tag}_*
