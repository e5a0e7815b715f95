set -x
mkdir -p gpurun_out
for p in 0 1 2; do
CSM_OPTIONS=bb_probe=$p python bench.py --no-cpu --no-single > gpurun_out/bench_probe$p.json 2> gpurun_out/bench_probe$p.err; tail -2 gpurun_out/bench_probe$p.err
python scripts/show_bench.py gpurun_out/bench_probe$p.json | grep -E "^value|^ms_per_step|^e2e|^warm|^repeats|^roofline"
done
CSM_OPTS=bb_probe=1 python scripts/exp_phases.py 256 refine 2>&1 | grep -E "dive|sum"
