set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python scripts/exp_bb.py 2>&1 | tail -6
python scripts/exp_cfg1.py 200 2>&1 | tail -7
python bench.py --steps 20 --warmup 3 --no-cpu --no-single > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "rc=$?"; tail -5 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
for k in ("value", "ms_per_step", "e2e", "phases", "gpu_launches"):
    print(k, d.get(k))
PY
