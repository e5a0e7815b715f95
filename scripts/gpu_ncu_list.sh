set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu --no-single > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-single > gpurun_out/ncu.log 2>&1
echo "rc=$?"
tail -3 gpurun_out/ncu.log
wc -l gpurun_out/launches.csv
