# closing run after the k_project work: smoke, both bench arms (timed by wall clock too), events of one step,
# ncu launch list of the bench and a full capture of k_project
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py > gpurun_out/bench_r2c_n1.json 2> gpurun_out/bench_r2c_n1.err ) 2>&1 | tail -4; tail -3 gpurun_out/bench_r2c_n1.err
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2c_ref.json 2> gpurun_out/bench_r2c_ref.err ) 2>&1 | tail -4
python scripts/exp_phases.py 256 refine > gpurun_out/phases_256_r2c.log 2>&1; tail -17 gpurun_out/phases_256_r2c.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-single"
$CMD > gpurun_out/plain_r2c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2c_launches.csv $CMD > gpurun_out/ncu_l_r2c.log 2>&1
echo "rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_project' --launch-skip 13 -c 1 \
    -o gpurun_out/r2c_project -f python scripts/exp_phases.py 256 refine > gpurun_out/ncu_r2c.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/r2c_project.ncu-rep --page raw --csv > gpurun_out/r2c_raw.csv 2>/dev/null; wc -l gpurun_out/r2c_raw.csv
