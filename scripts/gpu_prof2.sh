set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-single"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_bb_expand|k_bb_roots" -s 14 -c 7 -o gpurun_out/prof_bb2 -f $CMD > gpurun_out/ncu_bb2.log 2>&1
echo "rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_t3.csv $CMD > gpurun_out/ncu3.log 2>&1
echo "rc=$?"
