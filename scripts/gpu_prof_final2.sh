set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-single"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "rc=$?"
python scripts/exp_cfg4.py > gpurun_out/plain_cfg4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_window_tma -s 1 -c 1 -o gpurun_out/prof_wt -f python scripts/exp_cfg4.py > gpurun_out/ncu_wt.log 2>&1
echo "rc=$?"
tail -4 gpurun_out/plain_cfg4.log
