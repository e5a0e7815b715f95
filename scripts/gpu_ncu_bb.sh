set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu --no-single > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-single > gpurun_out/ncu.log 2>&1
echo "rc=$?"
python bench.py --steps 2 --warmup 1 --no-cpu --no-single > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_bb_score -s 14 -c 7 -o gpurun_out/prof_bb -f python bench.py --steps 2 --warmup 1 --no-cpu --no-single > gpurun_out/ncu2.log 2>&1
echo "rc=$?"
ls -la gpurun_out
