# A/B: sweep CTAs per SM (4 fill the register file; fewer leave room for the other kernels of steps in flight)
for c in 0 2 1 0 2; do
  CSM_OPTIONS="bb_sweep_ctas_per_sm=$c" python bench.py --no-cpu --no-single > gpurun_out/bench_ctas$c.json 2> gpurun_out/bench_ctas$c.err
  echo "cap $c"; python scripts/show_bench.py gpurun_out/bench_ctas$c.json 2>/dev/null | grep -E "^value|^ms_per_step|^e2e |^warm" | cut -c1-200
done
