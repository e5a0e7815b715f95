mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_pyramid_stream -s 3 -c 1 -o gpurun_out/prof_pyr2 -f python scripts/exp_phases.py 256 > gpurun_out/ncu_pyr2.log 2>&1
CSM_OPTS=pyramid_mode=3 ncu --set full --clock-control none --import-source on -k regex:k_pyramid_stream -s 3 -c 1 -o gpurun_out/prof_pyr1 -f python scripts/exp_phases.py 256 > gpurun_out/ncu_pyr1.log 2>&1
for n in 1 2; do ncu -i gpurun_out/prof_pyr$n.ncu-rep --page raw --csv > gpurun_out/prof_pyr${n}_raw.csv 2>/dev/null; done
ls -la gpurun_out/*.csv
