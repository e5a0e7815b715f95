# A/B: handles in flight x sweep CTAs per SM
for cfg in "3 2" "6 2" "8 2" "6 3" "4 2"; do
  set -- $cfg
  CSM_BENCH_LANES=$1 CSM_BENCH_SWEEP_CTAS=$2 python bench.py --no-cpu --no-single > gpurun_out/bench_l$1c$2.json 2> gpurun_out/bench_l$1c$2.err
  echo "lanes $1 ctas $2"; python scripts/show_bench.py gpurun_out/bench_l$1c$2.json 2>/dev/null | grep -E "^value|^warm" | cut -c1-30,150-215
done
