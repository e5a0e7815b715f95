set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -30
python bench.py --steps 20 --warmup 3 --no-single > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
for k in ("value", "ms_per_step", "e2e", "roofline", "phases", "gpu_launches", "clocks", "check", "cpu_baseline"):
    print(k, d.get(k))
PY
