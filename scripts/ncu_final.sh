# Dev tool (GPU): the round's ncu evidence from the bench command: launch list of the default step, then `--set full`
# captures of the step's kernels (C2), of the extraction kernels (C5) and of the wide-network kernels (C4).
tag=${1:-final}
o=gpurun_out
python bench.py --steps 16 --warmup 3 --no-also --no-cpu-baseline > $o/${tag}_plain_c2.json 2>$o/${tag}_plain_c2.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_c2.csv python bench.py --steps 16 --warmup 3 --no-also --no-cpu-baseline > $o/${tag}_ncu0.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:mlp_jet --launch-skip 8 --launch-count 2 -f -o $o/${tag}_U_c2 python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:mlp_plain --launch-skip 8 --launch-count 2 -f -o $o/${tag}_N_c2 python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_ncu2.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name "regex:gram_kernel|rfe_kernel|mlp_jet_fwd|mlp_plain_fwd" --launch-skip 33 --launch-count 4 -f -o $o/${tag}_extract_c5 python bench.py --config C5 --steps 2 --warmup 1 --no-cpu-baseline --no-also > $o/${tag}_ncu3.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:mlp_jet --launch-skip 4 --launch-count 2 -f -o $o/${tag}_U_c4 python bench.py --config C4 --points 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-also > $o/${tag}_ncu4.log 2>&1
ls -la $o/${tag}_*
