set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-single"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "rc=$?"
CMD2="python scripts/exp_phases.py 256 refine"
$CMD2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_bb_expand|k_bb_roots|k_pyramid_stream|k_project|k_finalize|k_refine" -s 33 -c 11 -o gpurun_out/prof_r1 -f $CMD2 > gpurun_out/ncu_f.log 2>&1
echo "rc=$?"
cat gpurun_out/plain2.log
