# round 2, after the incumbent dive: events of one step, full ncu capture of the sweep / dive / builder, launch list of the bench
set -x
mkdir -p gpurun_out
python scripts/exp_phases.py 256 refine > gpurun_out/phases_256_r2b.log 2>&1 && cat gpurun_out/phases_256_r2b.log
ncu --set full --clock-control none --import-source on -k 'regex:k_bbg_expand|k_pyramid_stream2|k_bbg_dive' --launch-skip 104 -c 10 \
    -o gpurun_out/r2b_sweep -f python scripts/exp_phases.py 256 refine > gpurun_out/ncu_r2b.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_r2b.log
ncu -i gpurun_out/r2b_sweep.ncu-rep --page raw --csv > gpurun_out/r2b_raw.csv 2>/dev/null; wc -l gpurun_out/r2b_raw.csv
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-single"
$CMD > gpurun_out/plain_r2b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2b_launches.csv $CMD > gpurun_out/ncu_l_r2b.log 2>&1
echo "rc=$?"
