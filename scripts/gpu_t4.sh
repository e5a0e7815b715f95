set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python scripts/exp_phases.py 64 | head -3
python scripts/exp_phases.py 128 | head -3
python scripts/exp_phases.py 256 | head -3
python scripts/exp_e2e2.py
