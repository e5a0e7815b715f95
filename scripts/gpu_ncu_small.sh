set -x
mkdir -p gpurun_out
for n in 64 128 256; do python scripts/exp_phases.py $n > gpurun_out/phases_$n.log 2>&1; cat gpurun_out/phases_$n.log; done
ncu --set full --clock-control none --import-source on -k regex:k_bb_expand -s 12 -c 6 -o gpurun_out/prof_bb64 -f python scripts/exp_phases.py 64 > gpurun_out/ncu_bb64.log 2>&1
echo rc=$?
ncu -i gpurun_out/prof_bb64.ncu-rep --page raw --csv > gpurun_out/prof_bb64_raw.csv 2>/dev/null
ls -la gpurun_out | head -30
