set -x
mkdir -p gpurun_out
python scripts/exp_phases.py 256 refine 2>&1 | grep -E "pyramid|sum"
python scripts/exp_phases.py 64 refine 2>&1 | grep -E "pyramid|sum"
python bench.py --no-cpu --no-single > gpurun_out/bench_pyr.json 2> gpurun_out/bench_pyr.err; tail -2 gpurun_out/bench_pyr.err
python scripts/show_bench.py gpurun_out/bench_pyr.json | grep -E "^value|^ms_per_step|^e2e |^pyr"
