set -x
mkdir -p gpurun_out
CMD="python scripts/exp_phases.py 256"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_bb_expand|k_bb_roots|k_project" -s 16 -c 8 -o gpurun_out/prof_bb4 -f $CMD > gpurun_out/ncu_bb4.log 2>&1
echo "rc=$?"
