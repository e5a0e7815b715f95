set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python scripts/exp_bb.py 2>&1 | tail -8
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-single"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_bb_score -s 14 -c 7 -o gpurun_out/prof_bb -f $CMD > gpurun_out/ncu_bb.log 2>&1
echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_pyramid_stream|k_project|k_finalize" -s 4 -c 4 -o gpurun_out/prof_misc -f $CMD > gpurun_out/ncu_misc.log 2>&1
echo "rc=$?"
ls -la gpurun_out
