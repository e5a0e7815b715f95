set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; tail -5 gpurun_out/bench_full.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_full.json"))
for k in ("value", "ms_per_step", "e2e", "e2e_dense", "roofline", "roofline_pyramid", "phases", "gpu_launches", "clocks", "check", "cpu_baseline", "single_scan"):
    print(k, d.get(k))
PY
