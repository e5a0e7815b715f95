set -x
mkdir -p gpurun_out
CMD="python scripts/exp_cfg4.py"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_window_tma" -s 1 -c 1 -o gpurun_out/prof_wt -f $CMD > gpurun_out/ncu_wt.log 2>&1
echo "rc=$?"
