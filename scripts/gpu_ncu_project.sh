# one full ncu capture of k_project / k_finalize / k_refine in a warmed-up 256-query step
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:k_project|k_theta_trig' --launch-skip 26 -c 2 \
    -o gpurun_out/r2c_project -f python scripts/exp_phases.py 256 refine > gpurun_out/ncu_r2c.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_r2c.log
ncu -i gpurun_out/r2c_project.ncu-rep --page raw --csv > gpurun_out/r2c_raw.csv 2>/dev/null; wc -l gpurun_out/r2c_raw.csv
ncu -i gpurun_out/r2c_project.ncu-rep --page source --csv -k regex:k_project > gpurun_out/r2c_project_source.csv 2>/dev/null; wc -l gpurun_out/r2c_project_source.csv
