set -x
mkdir -p gpurun_out
python scripts/exp_cfg1.py 200 > gpurun_out/cfg1.log 2>&1; cat gpurun_out/cfg1.log
python scripts/exp_cfg1.py 3 > gpurun_out/plain_cfg1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_cfg1.csv python scripts/exp_cfg1.py 3 > gpurun_out/ncu_cfg1.log 2>&1
echo rc=$?
