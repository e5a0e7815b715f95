set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for p in 0 1 2; do
CSM_OPTS=bb_probe=$p python scripts/exp_phases.py 256 refine > gpurun_out/phases_probe$p.log 2>&1; cat gpurun_out/phases_probe$p.log
done
for p in 0 2; do
CSM_OPTS=bb_probe=$p python scripts/exp_phases.py 32 refine > gpurun_out/phases_probe${p}_32.log 2>&1; cat gpurun_out/phases_probe${p}_32.log
done
