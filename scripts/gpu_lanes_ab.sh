set -x
mkdir -p gpurun_out
for l in 3 6 8; do
CSM_BENCH_LANES=$l python bench.py --no-cpu --no-single > gpurun_out/bench_lanes$l.json 2> gpurun_out/bench_lanes$l.err; tail -2 gpurun_out/bench_lanes$l.err
python scripts/show_bench.py gpurun_out/bench_lanes$l.json | grep -E "^value|^ms_per_step|^e2e |^host_issue"
done
